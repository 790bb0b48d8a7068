"""ADAP (pantheonrl/algos/adap): `ADAP`, `AdapPolicy`, `AdapAgent`, `SAMPLERS` with the reference's names and
signatures (trainer.py:127-130, 205-213; adap_learn.py:60-224; agent.py:21-151; util.py:42-131).

An AdapPolicy is the MlpPolicy whose two towers read cat(features, context) (policies.py:71-106); the context is
a [1, context_size] vector the policy carries, resampled at every episode end (adap_learn.py:452-455,
agent.py:146-150) or synchronised with another policy's (`latent_syncer`, trainer.py:210-213); every stored
observation row ends with the context it was chosen under.  ADAP.train is PPO.train plus
`context_loss_coeff * context_loss` (util.py:97-131).  All of it runs in libpantheon_b200.so: the forward with
context inputs is pth_policy_forward(context=), the update — context tiles included — is pth_ppo_update with
loss_kind PTH_LOSS_ADAP, the random draws come from pth_adap_draw (Philox) or, in the reference-RNG mode, from
torch's global generator at the points and in the order the reference draws them.

`AdapPolicyMult` (policies.py:134-283, `trainer.py ... ADAP_MULT`): the context scales a hidden layer instead of
joining the inputs — x = tanh(W0 f + b0); s = tanh(Ws x + bs) (64 -> 64 C); y_j = x_j + sum_c s[j C + c] ctx_c;
out = tanh(W1 y + b1) — same entry points with `adap_mult` set.

Host-driven flow only (n_envs = 1), like the reference.
"""
import numpy as np
import torch

from . import _lib, policy as pol, update as up
from .common.agents import OnPolicyAgent
from .ppo import PPO, DevicePolicy, HostStagedBuffer


def _torch_sampler(name, ctx_size, num):
    """SAMPLERS[name](ctx_size, num, torch=True) of pantheonrl/algos/adap/util.py:42-94, on torch's global generator."""
    if name == "l2":
        c = torch.rand(num, ctx_size) * 2 - 1
        return c / (torch.sum(c ** 2, dim=-1).reshape(num, 1)) ** (1 / 2)
    if name == "unit_square":
        return torch.rand(num, ctx_size) * 2 - 1
    if name == "positive_square":
        return torch.rand(num, ctx_size)
    if name == "categorical":
        c = torch.zeros(num, ctx_size)
        c[torch.arange(num), torch.randint(0, ctx_size, size=(num,))] = 1
        return c
    if name == "natural_numbers":
        return torch.randint(0, ctx_size, size=(num, 1)).float()
    raise KeyError(name)


def _numpy_sampler(name, ctx_size, num):
    """SAMPLERS[name](ctx_size, num, torch=False): the np.random branch of the same functions."""
    if name == "l2":
        c = np.random.rand(num, ctx_size) * 2 - 1
        return c / (np.sum(c ** 2, axis=-1).reshape(num, 1)) ** (1 / 2)
    if name == "unit_square":
        return np.random.rand(num, ctx_size) * 2 - 1
    if name == "positive_square":
        return np.random.rand(num, ctx_size)
    if name == "categorical":
        c = np.zeros((num, ctx_size))
        c[np.arange(num), np.random.randint(0, ctx_size, size=(num,))] = 1
        return c
    if name == "natural_numbers":
        return np.random.randint(0, ctx_size, size=(num, 1))
    raise KeyError(name)


def _make_sampler(name):
    def sampler(ctx_size, num, torch=False):  # the reference's signature (util.py:42-94)
        return _torch_sampler(name, ctx_size, num) if torch else _numpy_sampler(name, ctx_size, num)
    sampler.__name__ = f"get_{name}"
    return sampler


SAMPLERS = {k: _make_sampler(k) for k in ("l2", "unit_square", "positive_square", "categorical", "natural_numbers")}


class AdapPolicy:
    """Names the policy class in `ADAP(policy=AdapPolicy, ...)` (trainer.py:128, 206); the device-side object is
    AdapDevicePolicy."""


class AdapPolicyMult:
    """Names the policy class in `ADAP(policy=AdapPolicyMult, ...)` (trainer.py:130, 208): MultModel towers
    (pantheonrl/algos/adap/policies.py:134-283), the context scales a hidden layer."""


class AdapDevicePolicy(DevicePolicy):
    def __init__(self, space, observation_space, action_space, seed, device, rng_stream, rng, context_size, mult=False):
        self.context_size, self.mult = int(context_size), bool(mult)
        self.extra_inputs = 0 if self.mult else self.context_size  # AdapPolicyMult: no context columns in the first layers
        super().__init__(space, observation_space, action_space, seed, device, rng_stream, rng)
        self.context = torch.zeros(1, self.context_size)  # a CPU tensor, like the reference's
        self._ctx_dev = torch.zeros(1, self.context_size, device=device)

    def set_context(self, ctxt):
        self.context = torch.as_tensor(ctxt, dtype=torch.float32).reshape(1, self.context_size).cpu()
        self._ctx_dev.copy_(self.context)

    def get_context(self):
        return self.context

    def _context(self):
        return self._ctx_dev

    # ---- AdapPolicyMult: its own parameter layout
    def _init_flat(self, space, seed):
        return pol.init_flat_mult(space, seed, self.context_size) if self.mult else super()._init_flat(space, seed)

    def forward(self, obs, deterministic=False):
        if not self.mult:
            return super().forward(obs, deterministic)
        from . import ops
        self._stage_obs(obs)
        race = None
        if self.rng == "reference":
            race = torch.cat([torch.empty(1, n).exponential_(1) for n in self.space.heads], dim=1).to(self.device)
        ops.policy_forward(self.space, self.params, self._obs_dev, seed=self.seed, rng_stream=self.rng_stream,
                           tick=self.calls & 0xffffffff, slot=0, idx0=0, want=("action", "value", "logp"), race=race,
                           context=self._ctx_dev, adap_mult=True, out=self._res)
        self.calls += 1
        return self._fetch()

    def predict_values(self, obs):
        if not self.mult:
            return super().predict_values(obs)
        from . import ops
        self._stage_obs(obs)
        dummy = torch.zeros(1, 4, dtype=torch.uint8, device=self.device)
        ops.policy_forward(self.space, self.params, self._obs_dev, action_in=dummy, want=("value",), context=self._ctx_dev,
                           adap_mult=True, out={"value": self._res["value"]})
        return self._fetch()[1]

    def state_dict(self):
        if not self.mult:
            return super().state_dict()
        return pol.flat_to_state_dict_mult(self.space, self.params.cpu().numpy(), self.context_size)

    def load_state_dict(self, sd):
        if not self.mult:
            return super().load_state_dict(sd)
        self.params.copy_(torch.from_numpy(pol.state_dict_to_flat_mult(self.space, sd, self.context_size)))


class AdapBuffer(HostStagedBuffer):
    """Rows of observation ++ context (`full_obs_shape`, adap_learn.py:404-408): the context columns are kept in
    their own float32 array, which is what pth_ppo_update's d_context reads."""

    def __init__(self, n_steps, device, gamma, gae_lambda, box, row, context_size):
        super().__init__(n_steps, device, gamma, gae_lambda, box=box, row=row)
        self.C = int(context_size)
        self.h["ctx"] = np.zeros((n_steps, self.C), np.float32)
        self.d["ctx"] = torch.zeros(n_steps, self.C, device=device)

    def add(self, obs, action, reward, episode_start, value, log_prob):
        flat = np.asarray(obs, np.float64).reshape(-1)
        self.h["ctx"][self.pos] = flat[-self.C:]
        super().add(flat[:-self.C], action, reward, episode_start, value, log_prob)


class ADAP(PPO):
    """ADAP(policy=AdapPolicy, env=, ..., context_loss_coeff=0.1, context_size=3, num_context_samples=5,
    context_sampler="l2", num_state_samples=32) — adap_learn.py:60-224."""

    def __init__(self, policy=AdapPolicy, env=None, *args, context_loss_coeff=0.1, context_size=3,
                 num_context_samples=5, context_sampler="l2", num_state_samples=32, policy_kwargs=None, **kw):
        self.mult = policy is AdapPolicyMult or policy == "AdapPolicyMult" or bool(kw.pop("mult", False))
        if context_sampler not in up.ADAP_SAMPLERS:
            raise KeyError(context_sampler)
        if not 1 <= int(context_size) <= 8:
            raise ValueError("context_size must be 1..8")
        if context_sampler == "natural_numbers" and context_size != 1:
            raise ValueError("natural_numbers contexts have one column (util.py:85-94): context_size must be 1")
        self.context_loss_coeff, self.context_size = float(context_loss_coeff), int(context_size)
        self.num_context_samples, self.num_state_samples = int(num_context_samples), int(num_state_samples)
        self.context_sampler = context_sampler
        self._ctx_draws = 0  # Philox counters of the per-episode and the per-minibatch draws
        self._loss_draws = 0
        kw.pop("use_sde", None), kw.pop("sde_sample_freq", None), kw.pop("create_eval_env", None)
        kw.pop("_init_setup_model", None)
        for k in ("clip_range_vf", "target_kl"):
            if kw.pop(k, None) is not None:
                raise NotImplementedError(f"{k} is not supported")
        super().__init__("AdapPolicy", env, *args, **kw)
        if self.n_envs != 1:
            raise _lib.PthError("ADAP runs the host-driven flow (n_envs = 1), like the reference")
        self.full_obs_shape = None
        self.last_context_loss = None
        # adap_learn.py:212-217: the first context, drawn right after the policy is built
        self.policy.set_context(self._sample_context())

    @staticmethod
    def _policy_ok(policy):
        return policy == "AdapPolicy"

    def _make_policy(self, eff_seed, stream):
        return AdapDevicePolicy(self.space, self.observation_space, self.action_space, eff_seed, self.device, stream,
                                self.rng, self.context_size, self.mult)

    def _make_buffer(self, n_steps, gamma, gae_lambda):
        return AdapBuffer(n_steps, self.device, gamma, gae_lambda, self.space.obs_kind == _lib.PTH_OBS_BOX,
                          self.space.row_bytes, self.context_size)

    # ---------------------------------------------------------------- random draws
    def _sample_context(self):
        """SAMPLERS[self.context_sampler](ctx_size=, num=1, torch=True): a [1, C] CPU tensor."""
        if self.rng == "reference":
            return _torch_sampler(self.context_sampler, self.context_size, 1)
        _, d = up.adap_draw(1, 1, self.context_size, self.context_sampler, self.policy.seed,
                            self.policy.rng_stream + 0x10000, index0=self._ctx_draws, device=self.device)
        self._ctx_draws += 1
        return d.reshape(1, self.context_size).cpu()

    def _loss_draws_for(self, M):
        """Per minibatch, in the reference's order (util.py:106, 113-114): the sampled states, then K contexts."""
        n_mb = -(-M // self.batch_size)
        n, K, S, C = self.n_epochs * n_mb, self.num_context_samples, self.num_state_samples, self.context_size
        if self.rng != "reference":
            st, dr = up.adap_draw(n, K, C, self.context_sampler, self.policy.seed, self.policy.rng_stream + 0x20000,
                                  index0=self._loss_draws, S=S, n_mb=n_mb, M=M, batch_size=self.batch_size,
                                  device=self.device)
            self._loss_draws += n
            return st, dr
        st, dr = np.full((n, S), -1, np.int32), np.zeros((n, K, C), np.float32)
        for i in range(n):
            B = min(self.batch_size, M - (i % n_mb) * self.batch_size)
            idx = torch.randperm(B)[:S].numpy()
            st[i, :len(idx)] = idx
            for k in range(K):
                dr[i, k] = _torch_sampler(self.context_sampler, C, 1).numpy().reshape(-1)
        return torch.from_numpy(st).to(self.device), torch.from_numpy(dr).to(self.device)

    # ---------------------------------------------------------------- ADAP.train (adap_learn.py:229-347)
    def train(self):
        buf, M = self.rollout_buffer, self.rollout_buffer.T
        if self._ws is None:
            self._ws = up.UpdateWorkspace(self.space, M, self.batch_size, self.device, context_size=self.context_size,
                                          adap_mult=self.mult)
            self._perm = torch.empty(self.n_epochs, M, dtype=torch.int32, device=self.device)
            self._ctx_loss = torch.zeros(self.n_epochs * (-(-M // self.batch_size)), device=self.device)
        if self.rng == "reference":
            # RolloutBuffer.get draws one np.random.permutation at the start of every epoch; the context loss
            # draws from torch's generator: two generators, so the interleaving does not matter
            self._perm.copy_(torch.from_numpy(np.stack([np.random.permutation(M) for _ in range(self.n_epochs)])
                                              .astype(np.int32)))
        else:
            up.perm_feistel(M, self.n_epochs, self.policy.seed, self.policy.rng_stream + 1, epoch0=self._n_updates,
                            out=self._perm)
        extra = {}
        if self.context_loss_coeff != 0.0 and self.num_context_samples >= 2:
            st, dr = self._loss_draws_for(M)
            extra = dict(loss_kind=_lib.PTH_LOSS_ADAP, context_loss_coeff=self.context_loss_coeff, ctx_states=st,
                         ctx_draws=dr, ctx_loss=self._ctx_loss)
        d = buf.d
        self.last_stats = up.ppo_update(
            self.space, self.policy.params, self.adam_m, self.adam_v, self.adam_step, d["obs"], d["actions"],
            d["logp"], d["advantages"], d["returns"], self._perm, self.batch_size, self._ws,
            learning_rate=self.learning_rate, clip_range=self.clip_range, ent_coef=self.ent_coef,
            vf_coef=self.vf_coef, max_grad_norm=self.max_grad_norm, normalize_advantage=self.normalize_advantage,
            context=d["ctx"], adap_mult=self.mult, **extra)
        self.last_context_loss = self._ctx_loss if extra else None
        self.adam_step += self.n_epochs * (-(-M // self.batch_size))
        self._n_updates += self.n_epochs
        self._record_train(self.last_stats, d["values"], d["returns"], self._n_updates)

    def _record_train(self, stats, values, returns, n_updates):
        super()._record_train(stats, values, returns, n_updates)
        if self._logger.output_formats and self.last_context_loss is not None:
            # adap_learn.py:358: context_kl_divs is reset every epoch, the logged value is the LAST epoch's mean
            cl = self.last_context_loss.cpu().numpy()
            self._logger.record("train/context_kl_loss", float(cl[-max(1, len(cl) // max(1, self.n_epochs)):].mean()))

    # ---------------------------------------------------------------- ADAP.collect_rollouts (adap_learn.py:377-473)
    def _collect(self, env, buf):
        policy = self.policy
        buf.reset()
        for _ in range(self.n_steps):
            actions, values, log_probs = policy.forward(self._last_obs)
            new_obs, reward, done, _info = env.step(actions[0])
            self.num_timesteps += 1
            self._ep[0] += float(reward)
            self._ep[1] += 1
            row = np.concatenate((np.asarray(self._last_obs, np.float64).reshape(-1),
                                  policy.get_context().numpy().astype(np.float64).reshape(-1)))
            buf.add(row, actions, reward, self._last_start, values, log_probs)
            self._last_start = done
            if done:
                self.ep_info_buffer.append({"r": self._ep[0], "l": self._ep[1]})
                self._ep = [0.0, 0]
            self._last_obs = env.reset() if done else new_obs  # DummyVecEnv auto-reset (inside env.step)
            if done:  # ADAP CHANGE: resample context (adap_learn.py:452-455)
                policy.set_context(self._sample_context())
        # adap_learn.py:457-460 bootstraps with policy.forward (a sample is drawn and dropped), not predict_values
        _, last_values, _ = policy.forward(self._last_obs)
        buf.compute_returns_and_advantage(last_values, self._last_start)

    def _learn_on_device(self, *a, **k):
        raise _lib.PthError("ADAP runs the host-driven flow (n_envs = 1), like the reference")

    # ---------------------------------------------------------------- checkpoint
    _HYPER = PPO._HYPER + ("context_loss_coeff", "context_size", "num_context_samples", "context_sampler",
                           "num_state_samples", "mult")

    def _layout(self):
        """(tensor names, flat -> state dict, state dict -> flat) of this policy's parameter vector."""
        C = self.context_size
        if self.mult:
            return ([n for n, _ in pol.tensor_shapes_mult(self.space, C)],
                    lambda f: pol.flat_to_state_dict_mult(self.space, f, C),
                    lambda sd: pol.state_dict_to_flat_mult(self.space, sd, C))
        return ([n for n, _ in pol.tensor_shapes(self.space, C)], lambda f: pol.flat_to_state_dict(self.space, f, C),
                lambda sd: pol.state_dict_to_flat(self.space, sd, C))

    def save(self, path):
        from . import checkpoint as ck
        names, to_dict, _ = self._layout()
        return ck.save_zip(
            path, self.observation_space, self.action_space, {k: getattr(self, k) for k in self._HYPER},
            self.policy.state_dict(),
            ck.optimizer_state_dict(names, to_dict(self.adam_m.cpu().numpy()), to_dict(self.adam_v.cpu().numpy()),
                                    self.adam_step, self.learning_rate),
            {"num_timesteps": self.num_timesteps, "n_updates": self._n_updates, "adam_step": self.adam_step})

    @classmethod
    def load(cls, path, env=None, **kw):
        from . import checkpoint as ck
        c = ck.load_zip(path)
        if env is None:
            env = type("_Spaces", (), {"observation_space": c["observation_space"],
                                       "action_space": c["action_space"]})()
        m = cls(AdapPolicy, env, **{**c["hyper"], **kw})  # hyper carries `mult` for an AdapPolicyMult archive
        m.policy.load_state_dict(c["policy"])
        names, _, to_flat = m._layout()
        mom_m, mom_v = ck.adam_moments(names, c["optimizer"])
        if all(v is not None for v in mom_m.values()):
            m.adam_m.copy_(torch.from_numpy(to_flat(mom_m)))
            m.adam_v.copy_(torch.from_numpy(to_flat(mom_v)))
        m.adam_step, m._n_updates = c["counters"]["adam_step"], c["counters"]["n_updates"]
        m.num_timesteps = c["counters"]["num_timesteps"]
        return m


class AdapAgent(OnPolicyAgent):
    """An ADAP learner in the partner's seat (adap/agent.py:21-151).  `latent_syncer`: a policy whose context
    this agent copies before every decision (trainer.py:210-213, `--share-latent`); without one the agent draws
    a new context whenever its episode ends.

    Reference defect, not reproduced: agent.py:125-127 reshapes the (observation ++ context) row to the
    policy's plain observation shape, which raises for every space (the sizes differ by context_size), so the
    reference's AdapAgent cannot record; here the row keeps its context, as ADAP.collect_rollouts stores it."""

    def __init__(self, model, log_interval=None, tensorboard_log=None, tb_log_name="AdapAgent", latent_syncer=None):
        super().__init__(model, log_interval, tensorboard_log, tb_log_name)
        self.latent_syncer = latent_syncer

    def get_action(self, obs, record=True):
        if self.latent_syncer is not None:
            self.model.policy.set_context(self.latent_syncer.get_context())
        return super().get_action(obs, record)

    def _stored_obs(self, obs):
        return np.concatenate((np.asarray(obs, np.float64).reshape(-1),
                               self.model.policy.get_context().numpy().astype(np.float64).reshape(-1)))

    def update(self, reward, done):
        super().update(reward, done)
        if done and self.latent_syncer is None:
            self.model.policy.set_context(self.model._sample_context())
