"""Behaviour cloning behind the reference's ``BC`` surface.

Reference: pantheonrl/algos/bc.py — ``BC(observation_space, action_space, expert_data=,
ent_weight=1e-3, l2_weight=0.0)`` (:170-245), ``set_expert_data_loader`` (:247-268),
``_calculate_loss`` (:270-315), ``train(n_epochs= | n_batches=)`` (:317-357),
``save_policy`` / ``reconstruct_policy`` (:359-370, :77-91), and the ``BCShell`` holder
``trainer.py:152`` wraps a loaded policy in.  The loss

    -mean(log_prob(acts | obs)) - ent_weight * mean(entropy) + l2_weight * sum(theta^2) / 2

and torch's default Adam (lr 1e-3, eps 1e-8) run in the same fused update kernel as PPO
(``loss_kind = PTH_LOSS_BC``): every epoch is a keyed shuffle of the transitions cut into
batches of ``BC.DEFAULT_BATCH_SIZE``; all epochs of a ``train()`` call are ONE launch.

Deviations, stated: the policy is the 64-64 two-tower ``MlpPolicy`` of the rest of this
package (the reference defaults to ``FeedForward32Policy``, 32-32; its ``policy_class`` is a
parameter), and the per-epoch shuffle is the keyed Feistel permutation, not torch's
``DataLoader(shuffle=True)`` stream.
"""
import numpy as np
import torch

from . import _lib, logger as lg, update as up
from .common.trajsaver import TransitionsMinimal
from .ppo import DevicePolicy
from .spaces import to_pth_space

STREAM_SHUFFLE_BC = 6  # Philox stream of the per-epoch permutations


class BCShell:
    """What trainer.py:152 hands to StaticPolicyAgent: an object with ``.policy``."""

    def __init__(self, policy):
        self.policy = policy

    def get_policy(self):
        return self.policy


class BC:
    DEFAULT_BATCH_SIZE = 32

    def __init__(self, observation_space, action_space, *, expert_data=None, ent_weight=1e-3, l2_weight=0.0,
                 learning_rate=1e-3, adam_eps=1e-8, seed=0, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.PthError("BC needs a CUDA device: this path has no CPU implementation")
        self.observation_space, self.action_space = observation_space, action_space
        self.device = "cuda" if device in ("auto", "cuda") else device
        self.space = to_pth_space(observation_space, action_space)
        self.policy = DevicePolicy(self.space, observation_space, action_space, seed, self.device, _lib.STREAM_ALT)
        self.ent_weight, self.l2_weight = float(ent_weight), float(l2_weight)
        self.learning_rate, self.adam_eps, self.seed = float(learning_rate), float(adam_eps), int(seed)
        self.adam_m = torch.zeros_like(self.policy.params)
        self.adam_v = torch.zeros_like(self.policy.params)
        self.adam_step, self.epochs_done, self.batches_done = 0, 0, 0
        self.logger, self.last_stats = lg.Logger(), None
        self._data = None
        if expert_data is not None:
            self.set_expert_data_loader(expert_data)

    # ------------------------------------------------------------------ data
    def set_expert_data_loader(self, expert_data):
        """A ``TransitionsMinimal`` (or anything with ``.obs`` / ``.acts`` arrays): uploaded once,
        every batch of every epoch is gathered from it on the device."""
        if not isinstance(expert_data, TransitionsMinimal):
            expert_data = TransitionsMinimal(np.asarray(expert_data.obs), np.asarray(expert_data.acts))
        M = len(expert_data)
        if M == 0:
            raise ValueError("no expert transitions")
        box = self.space.obs_kind == _lib.PTH_OBS_BOX
        obs = np.zeros((M, _lib.PTH_OC_ROW), np.float32) if box else np.zeros((M, self.space.row_bytes), np.uint8)
        flat = np.asarray(expert_data.obs).reshape(M, -1)
        obs[:, :flat.shape[1]] = flat
        acts = np.zeros((M, 4), np.uint8)
        a = np.asarray(expert_data.acts).reshape(M, -1)
        acts[:, :a.shape[1]] = a
        d = self.device
        self._data = dict(M=M, obs=torch.from_numpy(obs).to(d), acts=torch.from_numpy(acts).to(d),
                          zeros=torch.zeros(M, dtype=torch.float32, device=d))
        self._ws = up.UpdateWorkspace(self.space, M, self.DEFAULT_BATCH_SIZE, d)

    # ------------------------------------------------------------------ training
    def _run(self, n_epochs, M_used, epoch0):
        d, bs = self._data, self.DEFAULT_BATCH_SIZE
        perm = up.perm_feistel(d["M"], n_epochs, self.seed, STREAM_SHUFFLE_BC, epoch0=epoch0, device=self.device)
        if M_used < d["M"]:  # the leading batches of one more shuffled epoch
            perm = perm[:, :M_used].contiguous()
        stats = up.ppo_update(self.space, self.policy.params, self.adam_m, self.adam_v, self.adam_step,
                              d["obs"], d["acts"], d["zeros"], d["zeros"], d["zeros"], perm, bs, self._ws, M=M_used,
                              learning_rate=self.learning_rate, ent_coef=self.ent_weight, vf_coef=0.0,
                              max_grad_norm=float("inf"), eps=self.adam_eps, normalize_advantage=False,
                              loss_kind=_lib.PTH_LOSS_BC, l2_weight=self.l2_weight)
        n_mb = -(-M_used // bs)
        self.adam_step += n_epochs * n_mb
        self.batches_done += n_epochs * n_mb
        return stats

    def train(self, *, n_epochs=None, n_batches=None, on_epoch_end=None, on_batch_end=None, log_interval=100):
        """Exactly one of ``n_epochs`` (full passes over the data) / ``n_batches`` (bc.py:317-357).
        The callbacks of the reference fire between device launches: ``on_epoch_end`` after every
        epoch when given (one launch per epoch then), ``on_batch_end`` is not supported."""
        if (n_epochs is None) == (n_batches is None):
            raise ValueError("Must provide exactly one of `n_epochs` and `n_batches` arguments.")
        if on_batch_end is not None:
            raise _lib.PthError("on_batch_end: batches of a launch run back to back on the device")
        if self._data is None:
            raise ValueError("set_expert_data_loader() first")
        M, bs = self._data["M"], self.DEFAULT_BATCH_SIZE
        per_epoch = -(-M // bs)
        full, rest = (n_epochs, 0) if n_epochs is not None else divmod(n_batches, per_epoch)
        chunks = []
        if on_epoch_end is not None:
            for _ in range(full):
                chunks.append(self._run(1, M, self.epochs_done))
                self.epochs_done += 1
                on_epoch_end()
        elif full:
            chunks.append(self._run(full, M, self.epochs_done))
            self.epochs_done += full
        if rest:
            chunks.append(self._run(1, min(M, rest * bs), self.epochs_done))
            self.epochs_done += 1
        self.last_stats = torch.cat(chunks) if chunks else None
        if self.last_stats is not None and self.logger.output_formats:
            s = self.last_stats.cpu().numpy()
            for i in range(0, len(s), max(1, log_interval)):
                for k, v in self.stats_row(s[i]).items():
                    self.logger.record(k, v)
                self.logger.dump(self.batches_done - len(s) + i)
        return self

    def stats_row(self, row):
        """One batch's statistics under the reference's names (bc.py:303-313)."""
        neglogp, entropy = float(row[0]), -float(row[2])
        ent_loss = -self.ent_weight * entropy
        return {"neglogp": neglogp, "entropy": entropy, "ent_loss": ent_loss, "prob_true_act": float(row[3]),
                "loss": neglogp + ent_loss, "grad_norm": float(row[6]), "batch_size": int(row[7])}

    # ------------------------------------------------------------------ policies on disk
    def save_policy(self, policy_path):
        """bc.py:359-365.  A torch file with SB3-named tensors + the spaces' descriptions."""
        from .checkpoint import _space_entry
        torch.save({"policy": self.policy.state_dict(), "observation_space": _space_entry(self.observation_space),
                    "action_space": _space_entry(self.action_space)}, policy_path)


def reconstruct_policy(policy_path, device="cuda"):
    """bc.py:77-91: load a policy saved by ``BC.save_policy`` (for FIXED / LOAD partners of type BC)."""
    from .checkpoint import space_from_entry
    blob = torch.load(policy_path, map_location="cpu", weights_only=True)
    obs_space, act_space = space_from_entry(blob["observation_space"]), space_from_entry(blob["action_space"])
    space = to_pth_space(obs_space, act_space)
    policy = DevicePolicy(space, obs_space, act_space, 0, "cuda" if device in ("auto", "cuda") else device,
                          _lib.STREAM_ALT)
    policy.load_state_dict(blob["policy"])
    return policy


__all__ = ["BC", "BCShell", "reconstruct_policy"]
