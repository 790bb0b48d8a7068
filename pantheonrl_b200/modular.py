"""ModularAlgorithm / ModularPolicy (pantheonrl/algos/modular) with the reference's names and signatures
(trainer.py:131-135: `ModularAlgorithm(policy=ModularPolicy, policy_kwargs=dict(num_partners=len(alt)), **kw)`).

ModularPolicy (modular/policies.py:23-396, defaults): the MlpPolicy plus one module per partner — two more 64-64
tanh towers, an action head and a value head, all reading the main policy tower's latent; with partner p in the
game, logits = main + module[p], value = main + module[p].  ModularAlgorithm.learn (modular/learn.py:353-404)
collects one rollout of n_steps per partner (`env.set_partnerid(p)` before every step) into that partner's
buffer, then train() runs, per partner, n_epochs of PPO on its buffer plus the marginal regulariser over ALL
modules (learn.py:221-351).  The forward is pth_policy_forward(num_partners, partner_idx), a partner phase of
train() is one launch of pth_ppo_update with loss_kind PTH_LOSS_MODULAR; GAE is pth_gae_f32.

Host-driven flow (n_envs = 1), like the reference.  One place where the stored rows differ from the reference's
(DESIGN.md 9): its `collect_rollouts` adds the FIRST row of every rollout with `episode_start = None`
(`self._last_dones = None`, :181, :211 — a NaN or a TypeError in SB3's buffer, depending on the numpy version);
here that row carries the done flag of the previous step, like SB3's `_last_episode_starts`.  GAE never reads
the first row's flag, so no number depends on it.  Kept as is: the bootstrap uses the value of the
LAST stored observation (:215, `values` of the loop), kept as is.  Adam skips the value modules of the partners
that are not being trained (they have no gradient; torch >= 2 `zero_grad` semantics — the pinned torch 1.13.1
would zero-fill them after their first use).  `baseline` / `nomain` policy options are not implemented.
"""
import math
from collections import deque

import numpy as np
import torch

from . import _lib, logger as lg, ops, policy as pol, update as up
from .ppo import PPO, DevicePolicy, HostStagedBuffer

HID = pol.HID


class ModularPolicy:
    """Names the policy class in `ModularAlgorithm(policy=ModularPolicy, ...)`; the device-side object is
    ModularDevicePolicy."""


def block_shapes(L):
    """(suffix, torch shape) of one partner module in flat order (oracle/pth_oracle_modular.inc: mod_block)."""
    return [("policy_net.0.weight", (HID, HID)), ("policy_net.0.bias", (HID,)), ("policy_net.2.weight", (HID, HID)),
            ("policy_net.2.bias", (HID,)), ("value_net.0.weight", (HID, HID)), ("value_net.0.bias", (HID,)),
            ("value_net.2.weight", (HID, HID)), ("value_net.2.bias", (HID,)), ("action.weight", (L, HID)),
            ("action.bias", (L,)), ("value.weight", (1, HID)), ("value.bias", (1,))]


def _block_name(p, suffix):
    if suffix.startswith("action."):
        return f"partner_action_net.{p}.{suffix[7:]}"
    if suffix.startswith("value."):
        return f"partner_value_net.{p}.{suffix[6:]}"
    return f"partner_mlp_extractor.{p}.{suffix}"


def init_flat(space, seed, num_partners):
    """ModularPolicy._build + do_init_weights (modular/policies.py:229-270): modules are created main first, then
    one (MlpExtractor, action_net, value_net) per partner; ONE orthogonal pass follows in the same order."""
    F, L = pol.feature_dim(space), sum(space.heads)
    if seed is not None:
        torch.manual_seed(int(seed))
    for fan_out, fan_in in ((HID, F), (HID, F), (HID, HID), (HID, HID), (L, HID), (1, HID)):
        torch.nn.Linear(fan_in, fan_out)  # creation draws the default init before orthogonal_ overwrites it
    for _ in range(num_partners):
        for fan_out, fan_in in ((HID, HID), (HID, HID), (HID, HID), (HID, HID), (L, HID), (1, HID)):
            torch.nn.Linear(fan_in, fan_out)
    g = math.sqrt(2)
    z = torch.zeros
    o = pol._ortho
    pi0, pi1, vf0, vf1 = o(HID, F, g), o(HID, HID, g), o(HID, F, g), o(HID, HID, g)
    parts = [pi0.t().contiguous(), z(HID), pi1, z(HID), vf0.t().contiguous(), z(HID), vf1, z(HID), o(L, HID, 0.01), z(L),
             o(1, HID, 1.0), z(1)]
    for _ in range(num_partners):
        parts += [o(HID, HID, g), z(HID), o(HID, HID, g), z(HID), o(HID, HID, g), z(HID), o(HID, HID, g), z(HID),
                  o(L, HID, 0.01), z(L), o(1, HID, 1.0), z(1)]
    return torch.cat([p.reshape(-1) for p in parts]).numpy().astype(np.float32)


class ModularDevicePolicy(DevicePolicy):
    def __init__(self, space, observation_space, action_space, seed, device, rng_stream, rng, num_partners):
        self.num_partners = int(num_partners)
        super().__init__(space, observation_space, action_space, seed, device, rng_stream, rng)

    def _init_flat(self, space, seed):
        return init_flat(space, seed, self.num_partners)

    def forward(self, obs, partner_idx=0, deterministic=False):
        """ModularPolicy.forward(obs, partner_idx) -> (actions, values, log_probs) (modular/policies.py:273-290)."""
        self._stage_obs(obs)
        race = None
        if self.rng == "reference":
            race = torch.cat([torch.empty(1, n).exponential_(1) for n in self.space.heads], dim=1).to(self.device)
        ops.policy_forward(self.space, self.params, self._obs_dev, seed=self.seed, rng_stream=self.rng_stream,
                           tick=self.calls & 0xffffffff, slot=0, idx0=0, want=("action", "value", "logp"),
                           race=race, num_partners=self.num_partners, partner_idx=int(partner_idx), out=self._res)
        self.calls += 1
        return self._fetch()

    def _names_shapes(self):
        out = pol.tensor_shapes(self.space)
        for p in range(self.num_partners):
            out += [(_block_name(p, sfx), shape) for sfx, shape in block_shapes(sum(self.space.heads))]
        return out

    def flat_to_dict(self, flat):
        flat = torch.as_tensor(np.asarray(flat, np.float32))
        out, o = {}, 0
        for i, (name, shape) in enumerate(self._names_shapes()):
            n = int(np.prod(shape))
            chunk = flat[o:o + n]
            if i in (0, 4):  # the main first layers are stored input-major
                chunk = chunk.reshape(shape[1], shape[0]).t().contiguous()
            out[name] = chunk.reshape(shape).clone()
            o += n
        return out

    def dict_to_flat(self, sd):
        parts = []
        for i, (name, shape) in enumerate(self._names_shapes()):
            t = torch.as_tensor(sd[name]).float().reshape(shape)
            if i in (0, 4):
                t = t.t().contiguous()
            parts.append(t.reshape(-1))
        return torch.cat(parts).numpy().astype(np.float32)

    def state_dict(self):
        return self.flat_to_dict(self.params.cpu().numpy())

    def load_state_dict(self, sd):
        self.params.copy_(torch.from_numpy(self.dict_to_flat(sd)))


class ModularAlgorithm(PPO):
    """ModularAlgorithm(policy=ModularPolicy, env=, policy_kwargs=dict(num_partners=N), marginal_reg_coef=0.0, ...)
    — modular/learn.py:21-135."""

    def __init__(self, policy=ModularPolicy, env=None, *args, policy_kwargs=None, marginal_reg_coef=0.0, **kw):
        pk = dict(policy_kwargs or {})
        self.num_partners = int(pk.pop("num_partners", kw.pop("num_partners", 1)))
        for k in ("baseline", "nomain"):
            if pk.pop(k, False):
                raise NotImplementedError(f"ModularPolicy({k}=True) is not implemented")
        pk.pop("partner_net_arch", None)
        if pk:
            raise NotImplementedError(f"policy_kwargs {sorted(pk)} are not supported")
        if not 1 <= self.num_partners <= 8:
            raise ValueError("num_partners must be 1..8")
        self.marginal_reg_coef = float(marginal_reg_coef)
        kw.pop("use_sde", None), kw.pop("sde_sample_freq", None), kw.pop("create_eval_env", None)
        kw.pop("_init_setup_model", None)
        for k in ("clip_range_vf", "target_kl"):
            if kw.pop(k, None) is not None:
                raise NotImplementedError(f"{k} is not supported")
        super().__init__("ModularPolicy", env, *args, **kw)
        if self.n_envs != 1:
            raise _lib.PthError("ModularAlgorithm runs the host-driven flow (n_envs = 1), like the reference")
        # one rollout buffer per partner (learn.py:133-141); `rollout_buffer` is that list, like the reference's
        self.rollout_buffer = [self._make_buffer(self.n_steps, self.gamma, self.gae_lambda)
                               for _ in range(self.num_partners)]
        self.vf_steps = [0] * self.num_partners  # optimiser steps of every partner's value modules
        self.last_marginal = None

    @staticmethod
    def _policy_ok(policy):
        return policy == "ModularPolicy"

    def _make_policy(self, eff_seed, stream):
        return ModularDevicePolicy(self.space, self.observation_space, self.action_space, eff_seed, self.device, stream,
                                   self.rng, self.num_partners)

    # ---------------------------------------------------------------- collect_rollouts (learn.py:155-218)
    def _collect_partner(self, env, buf, partner_idx):
        policy = self.policy
        buf.reset()
        values = None
        for _ in range(self.n_steps):
            actions, values, log_probs = policy.forward(self._last_obs, partner_idx)
            env.set_partnerid(partner_idx)  # learn.py:194, before every step
            new_obs, reward, done, _info = env.step(actions[0])
            self.num_timesteps += 1
            self._ep[0] += float(reward)
            self._ep[1] += 1
            buf.add(self._last_obs, actions, reward, self._last_start, values, log_probs)
            self._last_start = done
            if done:
                self.ep_info_buffer.append({"r": self._ep[0], "l": self._ep[1]})
                self._ep = [0.0, 0]
            self._last_obs = env.reset() if done else new_obs  # DummyVecEnv auto-reset
        # learn.py:215: the bootstrap value is the LAST forward's (the value of the last stored observation)
        buf.compute_returns_and_advantage(values, self._last_start)

    # ---------------------------------------------------------------- train (learn.py:221-351)
    def train(self):
        M, BS = self.n_steps, self.batch_size
        n_mb = -(-M // BS)
        n = self.n_epochs * n_mb
        if self._ws is None:
            self._ws = up.UpdateWorkspace(self.space, M, BS, self.device, num_partners=self.num_partners)
            self._perm = torch.empty(self.n_epochs, M, dtype=torch.int32, device=self.device)
            self._marg = torch.zeros(n, device=self.device)
        stats, margs = [], []
        for p, buf in enumerate(self.rollout_buffer):
            if self.rng == "reference":
                self._perm.copy_(torch.from_numpy(np.stack([np.random.permutation(M) for _ in range(self.n_epochs)])
                                                  .astype(np.int32)))
            else:
                up.perm_feistel(M, self.n_epochs, self.policy.seed, self.policy.rng_stream + 1,
                                epoch0=self._n_updates * self.num_partners + p * self.n_epochs, out=self._perm)
            d = buf.d
            st = up.ppo_update(
                self.space, self.policy.params, self.adam_m, self.adam_v, self.adam_step, d["obs"], d["actions"], d["logp"],
                d["advantages"], d["returns"], self._perm, BS, self._ws, learning_rate=self.learning_rate,
                clip_range=self.clip_range, ent_coef=self.ent_coef, vf_coef=self.vf_coef,
                max_grad_norm=self.max_grad_norm, normalize_advantage=True, loss_kind=_lib.PTH_LOSS_MODULAR,
                num_partners=self.num_partners, partner_idx=p, partner_vf_step=self.vf_steps[p],
                marginal_reg_coef=self.marginal_reg_coef, ctx_loss=self._marg)
            self.adam_step += n
            self.vf_steps[p] += n
            stats.append(st)
            margs.append(self._marg.clone())
        self._n_updates += self.n_epochs
        self.last_stats, self.last_marginal = torch.cat(stats), torch.cat(margs)
        if self._logger.output_formats:  # learn.py:337-339: the three scalars the reference records
            s = self.last_stats.cpu().numpy()
            self._logger.record("train/entropy_loss", float(s[:, 2].mean()))
            self._logger.record("train/policy_gradient_loss", float(s[:, 0].mean()))
            self._logger.record("train/value_loss", float(s[:, 1].mean()))

    # ---------------------------------------------------------------- learn (learn.py:353-404)
    def learn(self, total_timesteps, callback=None, log_interval=1, tb_log_name="OnPolicyAlgorithm",
              reset_num_timesteps=True, progress_bar=False):
        if callback is not None or progress_bar:
            raise NotImplementedError("learn(callback=, progress_bar=) are not supported")
        self._setup_learn(tb_log_name, reset_num_timesteps)
        env = self.env
        if len(env.partners[0]) < self.num_partners:
            raise _lib.PthError(f"ModularAlgorithm(num_partners={self.num_partners}) needs that many partners in the env")
        if self._last_obs is None:
            self._last_obs = env.reset()
            self._last_start = True
            self._ep = [0.0, 0]
        if self.ep_info_buffer is None:
            self.ep_info_buffer = deque(maxlen=100)
        target = self.num_timesteps + total_timesteps
        while self.num_timesteps < target:
            for p in range(self.num_partners):
                env.set_partnerid(p)
                self._collect_partner(env, self.rollout_buffer[p], p)
            self._iteration += 1
            if log_interval is not None and self._iteration % log_interval == 0 and self._logger.output_formats:
                eps = list(self.ep_info_buffer)
                self._record_rollout(self._iteration, lg.safe_mean(e["r"] for e in eps) if eps else None,
                                     lg.safe_mean(e["l"] for e in eps) if eps else None)
            self.train()
        return self

    def _learn_on_device(self, *a, **k):
        raise _lib.PthError("ModularAlgorithm runs the host-driven flow (n_envs = 1), like the reference")

    # ---------------------------------------------------------------- checkpoint
    _HYPER = PPO._HYPER + ("num_partners", "marginal_reg_coef")

    def save(self, path):
        from . import checkpoint as ck
        names = [n for n, _ in self.policy._names_shapes()]
        return ck.save_zip(
            path, self.observation_space, self.action_space, {k: getattr(self, k) for k in self._HYPER},
            self.policy.state_dict(),
            ck.optimizer_state_dict(names, self.policy.flat_to_dict(self.adam_m.cpu().numpy()),
                                    self.policy.flat_to_dict(self.adam_v.cpu().numpy()), self.adam_step,
                                    self.learning_rate),
            {"num_timesteps": self.num_timesteps, "n_updates": self._n_updates, "adam_step": self.adam_step,
             "vf_steps": list(self.vf_steps)})

    @classmethod
    def load(cls, path, env=None, **kw):
        from . import checkpoint as ck
        c = ck.load_zip(path)
        if env is None:
            env = type("_Spaces", (), {"observation_space": c["observation_space"],
                                       "action_space": c["action_space"]})()
        hyper = {**c["hyper"], **kw}
        m = cls(ModularPolicy, env, policy_kwargs={"num_partners": hyper.pop("num_partners", 1)}, **hyper)
        m.policy.load_state_dict(c["policy"])
        names = [n for n, _ in m.policy._names_shapes()]
        mom_m, mom_v = ck.adam_moments(names, c["optimizer"])
        if all(v is not None for v in mom_m.values()):
            m.adam_m.copy_(torch.from_numpy(m.policy.dict_to_flat(mom_m)))
            m.adam_v.copy_(torch.from_numpy(m.policy.dict_to_flat(mom_v)))
        m.adam_step, m._n_updates = c["counters"]["adam_step"], c["counters"]["n_updates"]
        m.num_timesteps = c["counters"]["num_timesteps"]
        m.vf_steps = [int(x) for x in c["counters"].get("vf_steps", [0] * m.num_partners)]
        return m
