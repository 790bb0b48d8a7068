"""PPO model object shaped like stable_baselines3.PPO as the reference uses it
(trainer.py:107-137, 196-203; agents.py:97-184): same keyword names, `.policy`,
`.rollout_buffer`, `.n_steps`, `.train()`, `.learn()`, `.save()` / `.load()`.

n_envs == 1: the reference's host-driven flow — one kernel call per decision
  (pth_policy_forward, B = 1), rows staged on the host and uploaded once per
  train(); GAE and PPO.train are pth_gae_f32 / pth_ppo_update.
n_envs  > 1: learn() hands the whole loop to the device engine (VecTrainer ->
  pth_rollout_run) for the built-in envs.
"""
import itertools
import time
from collections import deque

import numpy as np
import torch

from . import _lib, logger as lg, ops, policy as pol, update as up
from .common.agents import OnPolicyAgent, StaticPolicyAgent
from .spaces import to_pth_space


# Philox streams of the host-driven (n_envs = 1) flow.  The reference builds the ego and every
# PPO partner with the SAME seed (trainer.py:111-112, 198-199): same initial weights, but their
# action samples come from one shared torch generator that keeps advancing, so the two learners
# never see the same random numbers.  Counter-based Philox has no shared state to advance: every
# PPO constructed in this process therefore takes the next pair of streams (sampling, shuffle)
# above the five streams of the device engine.  Construction order defines the streams, so a
# script run twice gives the same trace.
FACADE_STREAM0 = 0x100
_instances = itertools.count()


def reset_stream_counter():
    """Start the per-process PPO instance count again (a fresh run inside one interpreter)."""
    global _instances
    _instances = itertools.count()


class DevicePolicy:
    """ActorCriticPolicy stand-in: parameters live on the device as one flat vector."""

    extra_inputs = 0  # AdapPolicy: context inputs behind the features

    def __init__(self, space, observation_space, action_space, seed, device, rng_stream, rng="philox"):
        self.space, self.observation_space, self.action_space = space, observation_space, action_space
        self.device, self.seed, self.rng_stream, self.rng = device, int(seed or 0), rng_stream, rng
        self.params = torch.from_numpy(self._init_flat(space, seed)).to(device)  # seed None: no re-seeding
        self.calls = 0
        self.box = space.obs_kind == _lib.PTH_OBS_BOX  # fp32 rows of 64 instead of 32 slot bytes
        self.row = space.row_bytes  # one-hot rows: 32 bytes, or 96 for frame-stacked observations
        self._obs_dev = (torch.zeros(1, _lib.PTH_OC_ROW, dtype=torch.float32, device=device) if self.box
                         else torch.zeros(1, self.row, dtype=torch.uint8, device=device))
        # one decision = one pinned upload of the observation row and ONE pinned download of
        # (action, value, log-prob): the three results share a 16-byte device buffer
        pin = str(device).startswith("cuda")
        self._obs_host = torch.zeros_like(self._obs_dev, device="cpu")
        self._res_dev = torch.zeros(16, dtype=torch.uint8, device=device)
        self._res_host = torch.zeros(16, dtype=torch.uint8)
        if pin:
            self._obs_host, self._res_host = self._obs_host.pin_memory(), self._res_host.pin_memory()
        self._obs_np = self._obs_host.numpy()
        self._res = {"action": self._res_dev[0:4].view(1, 4), "value": self._res_dev[4:8].view(torch.float32),
                     "logp": self._res_dev[8:12].view(torch.float32)}
        self.act_dim = space.n_heads

    def _init_flat(self, space, seed):
        return pol.init_flat(space, seed, self.extra_inputs)

    def _stage_obs(self, obs):
        flat = np.asarray(obs).reshape(-1)
        self._obs_np[0, :flat.size] = flat
        self._obs_np[0, flat.size:] = 0
        self._obs_dev.copy_(self._obs_host, non_blocking=True)

    def _fetch(self):
        """-> (actions [1, act_dim] int64, value tensor [1], log-prob tensor [1]) on the host, one copy + one sync."""
        self._res_host.copy_(self._res_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h = self._res_host.numpy()
        act = h[0:4].astype(np.int64)[None, :self.act_dim]
        if self.act_dim == 1 and getattr(self.action_space, "shape", ()) == ():
            act = act.reshape(1)  # Discrete: actions[0] is a scalar, like SB3
        vl = torch.from_numpy(h[4:12].view(np.float32).copy())
        return act, vl[0:1], vl[1:2]

    def _context(self):
        return None  # AdapPolicy: the [1, C] context on the device

    def forward(self, obs, deterministic=False):
        """obs -> (actions [1, act_dim] numpy, values tensor [1], log_probs tensor [1])."""
        self._stage_obs(obs)
        race = None
        if self.rng == "reference":
            # what Categorical.sample() draws from torch's default generator: torch.multinomial(probs, 1)
            # fills a tensor like probs with exponential_(1) and takes argmax(probs / q); one head after
            # the other (SB3 MultiCategoricalDistribution.sample)
            race = torch.cat([torch.empty(1, n).exponential_(1) for n in self.space.heads], dim=1).to(self.device)
        ops.policy_forward(self.space, self.params, self._obs_dev, seed=self.seed,
                           rng_stream=self.rng_stream, tick=self.calls & 0xffffffff, slot=0, idx0=0,
                           want=("action", "value", "logp"), race=race, context=self._context(), out=self._res)
        self.calls += 1
        return self._fetch()

    def predict_values(self, obs):
        """ActorCriticPolicy.predict_values: the value tower only — no sample is drawn."""
        self._stage_obs(obs)
        dummy = torch.zeros(1, 4, dtype=torch.uint8, device=self.device)
        ops.policy_forward(self.space, self.params, self._obs_dev, action_in=dummy, want=("value",),
                           context=self._context(), out={"value": self._res["value"]})
        return self._fetch()[1]

    def state_dict(self):
        return pol.flat_to_state_dict(self.space, self.params.cpu().numpy(), self.extra_inputs)

    def load_state_dict(self, sd):
        self.params.copy_(torch.from_numpy(pol.state_dict_to_flat(self.space, sd, self.extra_inputs)))


class HostStagedBuffer:
    """RolloutBuffer for the N = 1 flow: rows are staged on the host while the env is
    stepped from Python and uploaded once when GAE / train() run on the device."""

    def __init__(self, n_steps, device, gamma=0.99, gae_lambda=0.95, box=False, row=32):
        self.T, self.device, self.gamma, self.gae_lambda = n_steps, device, gamma, gae_lambda
        obs = np.zeros((n_steps, _lib.PTH_OC_ROW), np.float32) if box else np.zeros((n_steps, row), np.uint8)
        self.h = dict(obs=obs, actions=np.zeros((n_steps, 4), np.uint8),
                      rewards=np.zeros(n_steps, np.float32), values=np.zeros(n_steps, np.float32),
                      logp=np.zeros(n_steps, np.float32), episode_starts=np.zeros(n_steps, np.float32))
        self.d = {k: torch.zeros(v.shape, dtype=torch.from_numpy(v).dtype, device=device) for k, v in self.h.items()}
        self.d["advantages"] = torch.zeros(n_steps, 1, device=device)
        self.d["returns"] = torch.zeros(n_steps, 1, device=device)
        self.reset()

    def reset(self):
        self.pos, self.full = 0, False

    def add(self, obs, action, reward, episode_start, value, log_prob):
        i = self.pos
        flat = np.asarray(obs).reshape(-1)
        self.h["obs"][i] = 0
        self.h["obs"][i, :flat.size] = flat
        a = np.asarray(action).reshape(-1)
        self.h["actions"][i] = 0
        self.h["actions"][i, :a.size] = a
        self.h["rewards"][i] = reward
        self.h["episode_starts"][i] = float(episode_start)
        self.h["values"][i] = float(value.item())
        self.h["logp"][i] = float(log_prob.item())
        self.pos += 1
        self.full = self.pos == self.T

    def add_reward(self, reward):
        # agents.py:198 `rewards[pos - 1] += reward`; with pos == 0 the reference hits
        # row -1 of a just-reset buffer, which the next add() overwrites: the reward is lost.
        if self.pos > 0:
            self.h["rewards"][self.pos - 1] += np.float32(reward)

    def upload(self):
        for k, v in self.h.items():
            self.d[k].copy_(torch.from_numpy(v))

    def compute_returns_and_advantage(self, last_values, dones, gamma=None, gae_lambda=None):
        gamma = self.gamma if gamma is None else gamma
        gae_lambda = self.gae_lambda if gae_lambda is None else gae_lambda
        self.upload()
        T = self.T
        lv = last_values.reshape(1).to(self.device).float()
        dn = torch.tensor([float(dones)], device=self.device)
        ops.gae(self.d["rewards"].view(T, 1), self.d["values"].view(T, 1),
                self.d["episode_starts"].view(T, 1), lv, dn, gamma, gae_lambda,
                out=(self.d["advantages"], self.d["returns"]))


def collect_rollouts_single_env(algo, env, policy, buf, n_steps):
    """SB3 ``OnPolicyAlgorithm.collect_rollouts`` for ONE env behind ``DummyVecEnv`` + ``Monitor``
    (in-tree copy: pantheonrl/algos/adap/adap_learn.py:377-473): every step stores the observation the
    action was chosen on together with the PREVIOUS step's done flag as ``episode_start``; a finished
    episode is reset at once and the reset observation is what the next action sees; the rollout
    bootstraps from the value of the observation AFTER the last step and that step's done flag.
    ``algo`` carries the state that survives between rollouts (``_last_obs``, ``_last_start``,
    ``num_timesteps``, the Monitor's running episode ``_ep`` and ``ep_info_buffer``).  Plain host code:
    pinned against the in-tree copy by tests/test_collect_rollouts_cpu.py with stand-in policy / buffer."""
    buf.reset()
    for _ in range(n_steps):
        actions, values, log_probs = policy.forward(algo._last_obs)
        new_obs, reward, done, _info = env.step(actions[0])
        algo.num_timesteps += 1
        algo._ep[0] += float(reward)
        algo._ep[1] += 1
        buf.add(algo._last_obs, actions, reward, algo._last_start, values, log_probs)
        algo._last_start = done
        if done:
            algo.ep_info_buffer.append({"r": algo._ep[0], "l": algo._ep[1]})
            algo._ep = [0.0, 0]
        algo._last_obs = env.reset() if done else new_obs  # DummyVecEnv auto-reset
    last_values = policy.predict_values(algo._last_obs)
    buf.compute_returns_and_advantage(last_values, algo._last_start)


class PPO:
    def __init__(self, policy="MlpPolicy", env=None, learning_rate=3e-4, n_steps=2048, batch_size=64,
                 n_epochs=10, gamma=0.99, gae_lambda=0.95, clip_range=0.2, normalize_advantage=True,
                 ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, tensorboard_log=None, verbose=0, seed=None,
                 device="cuda", n_envs=1, n_minibatches=0, rng=None):
        if not self._policy_ok(policy):
            raise ValueError("only 'MlpPolicy' (SB3 default 64-64 tanh towers) is implemented")
        if not torch.cuda.is_available():
            raise _lib.PthError("PPO needs a CUDA device: this path has no CPU implementation")
        from .compat import unwrap_env
        env = unwrap_env(env)  # DummyVecEnv([lambda: Monitor(env)]) -> env (trainer.py:119)
        self.env, self.device = env, ("cuda" if device in ("auto", "cuda") else device)
        self.learning_rate, self.n_steps, self.batch_size, self.n_epochs = learning_rate, n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda, self.clip_range = gamma, gae_lambda, clip_range
        self.normalize_advantage, self.ent_coef, self.vf_coef = normalize_advantage, ent_coef, vf_coef
        self.max_grad_norm, self.verbose, self.seed = max_grad_norm, verbose, seed
        self.tensorboard_log, self.n_envs, self.n_minibatches = tensorboard_log, int(n_envs), n_minibatches
        self.observation_space, self.action_space = env.observation_space, env.action_space
        self.space = to_pth_space(self.observation_space, self.action_space)
        from .rng_mode import get_rng_mode
        self.rng = rng or get_rng_mode()
        if self.rng == "reference" and seed is not None:
            # SB3 set_random_seed(seed) (BaseAlgorithm._setup_model): every constructor re-seeds all three
            # global generators (torch is re-seeded by init_flat below, which then draws the weights)
            import random
            random.seed(seed)
            np.random.seed(seed)
        stream = FACADE_STREAM0 + 2 * next(_instances)  # its shuffle stream is stream + 1
        # seed=None: SB3 leaves the global generators alone, so two unseeded learners differ; here the
        # effective seed comes from numpy's global stream (np.random.seed(...) still pins a whole run)
        if seed is not None:
            eff_seed = seed
        elif self.rng == "reference":
            eff_seed = None  # no generator is touched: the weights come from torch's current state
        else:
            eff_seed = int(np.random.randint(1, 2 ** 31 - 1))
        self.policy = self._make_policy(eff_seed, stream)
        if self.space.obs_kind == _lib.PTH_OBS_BOX and self.space.obs_len > _lib.PTH_OC_ROW:
            raise _lib.PthError("Box observations wider than 64 are not supported")
        self.rollout_buffer = self._make_buffer(n_steps, gamma, gae_lambda)
        self.adam_m = torch.zeros_like(self.policy.params)
        self.adam_v = torch.zeros_like(self.policy.params)
        self.adam_step, self._n_updates, self.num_timesteps = 0, 0, 0
        self.use_sde, self.sde_sample_freq = False, -1
        self.last_stats = None
        self.ep_info_buffer = None
        self._logger, self._custom_logger = lg.Logger(), False
        self._num_timesteps_at_start, self.start_time, self._iteration = 0, time.time(), 0
        self._ws = None
        self._last_obs = None
        self._trainer = None

    # ---------------------------------------------------------------- what a subclass (ADAP) swaps
    @staticmethod
    def _policy_ok(policy):
        return policy == "MlpPolicy"

    def _make_policy(self, eff_seed, stream):
        return DevicePolicy(self.space, self.observation_space, self.action_space, eff_seed, self.device, stream,
                            self.rng)

    def _make_buffer(self, n_steps, gamma, gae_lambda):
        return HostStagedBuffer(n_steps, self.device, gamma, gae_lambda,
                                box=self.space.obs_kind == _lib.PTH_OBS_BOX, row=self.space.row_bytes)

    # ---------------------------------------------------------------- logging (SB3 Logger surface)
    @property
    def logger(self):
        return self._logger

    def set_logger(self, logger):
        """BaseAlgorithm.set_logger: a logger set by the caller survives learn()."""
        self._logger, self._custom_logger = logger, True

    def _record_train(self, stats, values, returns, n_updates):
        """The scalars SB3's PPO.train records (adap_learn.py:354-371): means over all
        minibatches, the last minibatch's loss, explained variance of the buffer."""
        if not self._logger.output_formats:
            return  # nothing would read them: skip the device -> host copy
        s = stats.cpu().numpy()
        r = self._logger.record
        r("train/entropy_loss", float(s[:, 2].mean()))
        r("train/policy_gradient_loss", float(s[:, 0].mean()))
        r("train/value_loss", float(s[:, 1].mean()))
        # SB3 resets approx_kl_divs at the start of every epoch: the logged value is the LAST epoch's mean
        # (adap_learn.py:250, 357; pinned by tests/test_oracle_sb3_intree.py)
        r("train/approx_kl", float(s[-max(1, len(s) // max(1, self.n_epochs)):, 3].mean()))
        r("train/clip_fraction", float(s[:, 4].mean()))
        r("train/loss", float(s[-1, 5]))
        r("train/explained_variance", lg.explained_variance(values, returns))
        r("train/n_updates", n_updates, exclude="tensorboard")
        r("train/clip_range", self.clip_range)

    def _record_rollout(self, iteration, ep_rew_mean, ep_len_mean):
        """OnPolicyAlgorithm.learn's per-iteration record + dump (adap_learn.py:487-500)."""
        elapsed = max(time.time() - self.start_time, 1e-9)
        r = self._logger.record
        r("time/iterations", iteration, exclude="tensorboard")
        if ep_rew_mean is not None:
            r("rollout/ep_rew_mean", ep_rew_mean)
            r("rollout/ep_len_mean", ep_len_mean)
        r("time/fps", int((self.num_timesteps - self._num_timesteps_at_start) / elapsed))
        r("time/time_elapsed", int(elapsed), exclude="tensorboard")
        r("time/total_timesteps", self.num_timesteps, exclude="tensorboard")
        self._logger.dump(step=self.num_timesteps)

    def _setup_learn(self, tb_log_name, reset_num_timesteps=True):
        """BaseAlgorithm._setup_learn: a learn() call counts its timesteps from 0 unless the caller
        continues a run with reset_num_timesteps=False (optimizer state and update counters always
        continue: they belong to the model, not to the call)."""
        if reset_num_timesteps:
            self.num_timesteps = 0
        if not self._custom_logger:
            self._logger = lg.configure_logger(self.verbose, self.tensorboard_log, tb_log_name)
        self.start_time, self._num_timesteps_at_start = time.time(), self.num_timesteps

    # ---------------------------------------------------------------- SB3 PPO.train
    def train(self):
        buf, M = self.rollout_buffer, self.rollout_buffer.T
        if self._ws is None:
            self._ws = up.UpdateWorkspace(self.space, M, self.batch_size, self.device)
            self._perm = torch.empty(self.n_epochs, M, dtype=torch.int32, device=self.device)
        if self.rng == "reference":  # SB3 RolloutBuffer.get: one np.random.permutation per epoch
            self._perm.copy_(torch.from_numpy(np.stack([np.random.permutation(M) for _ in range(self.n_epochs)])
                                              .astype(np.int32)))
        else:
            up.perm_feistel(M, self.n_epochs, self.policy.seed, self.policy.rng_stream + 1, epoch0=self._n_updates,
                            out=self._perm)
        d = buf.d
        self.last_stats = up.ppo_update(
            self.space, self.policy.params, self.adam_m, self.adam_v, self.adam_step, d["obs"], d["actions"],
            d["logp"], d["advantages"], d["returns"], self._perm, self.batch_size, self._ws,
            learning_rate=self.learning_rate, clip_range=self.clip_range, ent_coef=self.ent_coef,
            vf_coef=self.vf_coef, max_grad_norm=self.max_grad_norm, normalize_advantage=self.normalize_advantage)
        self.adam_step += self.n_epochs * (-(-M // self.batch_size))
        self._n_updates += self.n_epochs
        self._record_train(self.last_stats, d["values"], d["returns"], self._n_updates)

    # ---------------------------------------------------------------- learn
    def learn(self, total_timesteps, callback=None, log_interval=1, tb_log_name="PPO", reset_num_timesteps=True,
              progress_bar=False):
        if callback is not None or progress_bar:
            raise NotImplementedError("learn(callback=, progress_bar=) are not supported")
        self._setup_learn(tb_log_name, reset_num_timesteps)
        if self.n_envs > 1:
            return self._learn_on_device(total_timesteps, log_interval)
        env, buf = self.env, self.rollout_buffer
        if self._last_obs is None:
            self._last_obs = env.reset()
            self._last_start = True
            self._ep = [0.0, 0]  # Monitor: reward and length of the ego's running episode
        if self.ep_info_buffer is None:
            self.ep_info_buffer = deque(maxlen=100)
        target = self.num_timesteps + total_timesteps
        while self.num_timesteps < target:
            self._collect(env, buf)
            self._iteration += 1
            if log_interval is not None and self._iteration % log_interval == 0 and self._logger.output_formats:
                eps = list(self.ep_info_buffer)
                self._record_rollout(self._iteration, lg.safe_mean(e["r"] for e in eps) if eps else None,
                                     lg.safe_mean(e["l"] for e in eps) if eps else None)
            self.train()
        return self

    def _collect(self, env, buf):
        collect_rollouts_single_env(self, env, self.policy, buf, self.n_steps)

    def _learn_on_device(self, total_timesteps, log_interval=1):
        from .engine import PPOConfig, VecTrainer
        kind = getattr(self.env, "device_kind", None)
        if kind is None:
            raise _lib.PthError("n_envs > 1 needs a built-in env with a device twin "
                                "(RPS-v0, LiarsDice-v0, OvercookedMultiEnv-v0)")
        plist = self.env.partners[0]
        if len(plist) < 1:
            raise _lib.PthError("the on-device loop needs a partner: env.add_partner_agent(...)")
        if len(plist) > 1:
            return self._learn_on_device_partner_set(total_timesteps, log_interval, plist, kind)
        partner = plist[0]
        if self._trainer is None:
            mk = lambda m: PPOConfig(learning_rate=m.learning_rate, n_steps=m.n_steps, batch_size=m.batch_size,  # noqa: E731
                                     n_epochs=m.n_epochs, gamma=m.gamma, gae_lambda=m.gae_lambda,
                                     clip_range=m.clip_range, normalize_advantage=m.normalize_advantage,
                                     ent_coef=m.ent_coef, vf_coef=m.vf_coef, max_grad_norm=m.max_grad_norm,
                                     n_minibatches=m.n_minibatches or 32)
            if isinstance(partner, OnPolicyAgent):
                mode, alt_cfg = "ppo", mk(partner.model)
                if partner.model.n_steps != self.n_steps:
                    import warnings
                    warnings.warn("on-device loop: the partner trains at the ego's rollout boundaries; its "
                                  f"n_steps={partner.model.n_steps} is not used (ego n_steps={self.n_steps})")
            elif isinstance(partner, StaticPolicyAgent) and partner.policy is self.policy:
                mode, alt_cfg = "selfplay", None
            else:
                raise _lib.PthError("on-device partners: OnPolicyAgent(PPO) or StaticPolicyAgent(ego.policy)")
            self._trainer = VecTrainer(kind, self.n_envs, mk(self), alt_cfg, seed=self.policy.seed, partner=mode,
                                       probegostart=getattr(self.env, "probegostart", 0.5), device=self.device,
                                       **({"layout": self.env.layout_name, "ego_agent_idx": self.env.ego_agent_idx,
                                           "horizon": self.env.layout.horizon} if kind == "overcooked" else {}))
            self._trainer.ego.params = self.policy.params          # share storage with the facade objects
            self._trainer.ego.adam_m, self._trainer.ego.adam_v = self.adam_m, self.adam_v
            # counters the models bring along (PPO.load, or earlier n_envs = 1 training): Adam's bias
            # correction and the shuffle keys continue, and the env / sampling streams start at fresh ticks
            tr0 = self._trainer
            tr0.ego.adam_step, tr0.ego.n_updates = self.adam_step, self._n_updates
            tr0.tick_base = (self._n_updates * self.n_steps) & 0xffffffff
            if mode == "ppo":
                tr0.alt.params = partner.model.policy.params
                tr0.alt.adam_m, tr0.alt.adam_v = partner.model.adam_m, partner.model.adam_v
                tr0.alt.adam_step, tr0.alt.n_updates = partner.model.adam_step, partner.model._n_updates
                tr0.partner_decisions = partner.num_timesteps
            self._ep_prev = np.zeros(4)
        tr = self._trainer
        tr.num_timesteps = self.num_timesteps
        target = tr.num_timesteps + total_timesteps
        partner_model = partner.model if isinstance(partner, OnPolicyAgent) else None
        prev = self._ep_prev  # device episode counters are cumulative: keep the last reading across learn() calls
        while tr.num_timesteps < target:
            tr.collect()
            tr.compute_gae()
            self.num_timesteps = tr.num_timesteps
            self._iteration += 1
            log_now = log_interval is not None and self._iteration % log_interval == 0
            if log_now and self._logger.output_formats:
                # episodes finished during this rollout (device counters: episodes, reward sum, length sum)
                e = tr.carry.ep_stats.cpu().numpy().astype(np.float64)
                d, prev = e - prev, e
                self._ep_prev = prev
                n = max(d[0], 1.0)
                self._record_rollout(self._iteration, float(d[1] / n), float(d[2] / n))
            tr.train()
            self._n_updates = tr.ego.n_updates
            self.adam_step = tr.ego.adam_step
            self._record_train(tr.ego.last_stats, tr.ego_buf.values, tr.ego_buf.returns, tr.ego.n_updates)
            if partner_model is not None:
                partner_model._n_updates, partner_model.adam_step = tr.alt.n_updates, tr.alt.adam_step
                partner.num_timesteps = tr.partner_decisions
                if log_now and partner_model.logger.output_formats:
                    m = tr.alt_buf.count.clamp(max=tr.alt_buf.Tcap)
                    mask = torch.arange(tr.alt_buf.Tcap, device=m.device)[:, None] < m[None, :]
                    partner_model._record_train(tr.alt.last_stats, tr.alt_buf.values[mask],
                                                tr.alt_buf.returns[mask], tr.alt.n_updates)
                    partner_model.logger.record("name", partner.name, exclude="tensorboard")
                    partner_model.logger.record("time/total_timesteps", partner.num_timesteps, exclude="tensorboard")
                    partner_model.logger.dump(step=partner.num_timesteps)
        self.last_stats = tr.ego.last_stats
        return self

    def _learn_on_device_partner_set(self, total_timesteps, log_interval, plist, kind):
        """Several partners added to the env (trainer.py:216-228): the device engine pairs each with its
        own share of the envs (partner_set.PartnerSetTrainer) instead of drawing one per episode."""
        from .engine import PPOConfig
        from .partner_set import PartnerSetTrainer
        if not all(isinstance(a, OnPolicyAgent) for a in plist):
            raise _lib.PthError("on-device partner sets are made of OnPolicyAgent(PPO) learners")
        if self.n_envs % len(plist):
            raise _lib.PthError(f"n_envs={self.n_envs} must be a multiple of the {len(plist)} partners")
        if self._trainer is None:
            mk = lambda m: PPOConfig(learning_rate=m.learning_rate, n_steps=m.n_steps, batch_size=m.batch_size,  # noqa: E731
                                     n_epochs=m.n_epochs, gamma=m.gamma, gae_lambda=m.gae_lambda,
                                     clip_range=m.clip_range, normalize_advantage=m.normalize_advantage,
                                     ent_coef=m.ent_coef, vf_coef=m.vf_coef, max_grad_norm=m.max_grad_norm,
                                     n_minibatches=m.n_minibatches or 32)
            tr = PartnerSetTrainer(kind, self.n_envs, mk(self), mk(plist[0].model), partners_per_gpu=len(plist),
                                   seed=self.policy.seed, probegostart=getattr(self.env, "probegostart", 0.5),
                                   device=self.device,
                                   **({"layout": self.env.layout_name, "ego_agent_idx": self.env.ego_agent_idx,
                                       "horizon": self.env.layout.horizon} if kind == "overcooked" else {}))
            tr.ego.params, tr.ego.adam_m, tr.ego.adam_v = self.policy.params, self.adam_m, self.adam_v
            tr.ego.adam_step, tr.ego.n_updates = self.adam_step, self._n_updates
            tr.tick_base = (self._n_updates * self.n_steps) & 0xffffffff
            for ln, agent in zip(tr.lanes, plist):
                m = agent.model
                ln.learner.params, ln.learner.adam_m, ln.learner.adam_v = m.policy.params, m.adam_m, m.adam_v
                ln.learner.adam_step, ln.learner.n_updates = m.adam_step, m._n_updates
            self._trainer, self._ep_prev = tr, np.zeros(4)
        tr = self._trainer
        tr.num_timesteps = self.num_timesteps
        target = tr.num_timesteps + total_timesteps
        while tr.num_timesteps < target:
            tr.collect()
            tr.compute_gae()
            self.num_timesteps = tr.num_timesteps
            self._iteration += 1
            if log_interval is not None and self._iteration % log_interval == 0 and self._logger.output_formats:
                e = tr.carry.ep_stats.cpu().numpy().astype(np.float64)
                d, self._ep_prev = e - self._ep_prev, e
                n = max(d[0], 1.0)
                self._record_rollout(self._iteration, float(d[1] / n), float(d[2] / n))
            tr.train()
            self._n_updates, self.adam_step = tr.ego.n_updates, tr.ego.adam_step
            self._record_train(tr.ego.last_stats, tr.ego_buf.values, tr.ego_buf.returns, tr.ego.n_updates)
            for ln, agent in zip(tr.lanes, plist):
                agent.model._n_updates, agent.model.adam_step = ln.learner.n_updates, ln.learner.adam_step
                agent.num_timesteps += ln.M
        self.last_stats = tr.ego.last_stats
        return self

    # ---------------------------------------------------------------- checkpoint
    _HYPER = ("learning_rate", "n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "clip_range",
              "normalize_advantage", "ent_coef", "vf_coef", "max_grad_norm", "seed", "n_envs", "verbose",
              "n_minibatches")

    def save(self, path):
        """SB3-1.7.0-shaped zip (`data` JSON + policy.pth + policy.optimizer.pth): what
        trainer.py:419-432 writes and trainer.py:146-151 / PPO.load read."""
        from . import checkpoint as ck
        names = [n for n, _ in pol.tensor_shapes(self.space)]
        return ck.save_zip(
            path, self.observation_space, self.action_space, {k: getattr(self, k) for k in self._HYPER},
            self.policy.state_dict(),
            ck.optimizer_state_dict(names, pol.flat_to_state_dict(self.space, self.adam_m.cpu().numpy()),
                                    pol.flat_to_state_dict(self.space, self.adam_v.cpu().numpy()),
                                    self.adam_step, self.learning_rate),
            {"num_timesteps": self.num_timesteps, "n_updates": self._n_updates, "adam_step": self.adam_step})

    @classmethod
    def load(cls, path, env=None, **kw):
        """PPO.load(path) as trainer.py:149 calls it (no env: the spaces come from the
        archive) or with an env / keyword overrides like SB3's."""
        from . import checkpoint as ck
        c = ck.load_zip(path)
        if env is None:
            env = type("_Spaces", (), {"observation_space": c["observation_space"],
                                       "action_space": c["action_space"]})()
        m = cls("MlpPolicy", env, **{**c["hyper"], **kw})
        m.policy.load_state_dict(c["policy"])
        names = [n for n, _ in pol.tensor_shapes(m.space)]
        mom_m, mom_v = ck.adam_moments(names, c["optimizer"])
        if all(v is not None for v in mom_m.values()):
            m.adam_m.copy_(torch.from_numpy(pol.state_dict_to_flat(m.space, mom_m)))
            m.adam_v.copy_(torch.from_numpy(pol.state_dict_to_flat(m.space, mom_v)))
        m.adam_step, m._n_updates = c["counters"]["adam_step"], c["counters"]["n_updates"]
        m.num_timesteps = c["counters"]["num_timesteps"]
        return m

    def set_env(self, env):
        """trainer.py:119-123 (`model.set_env(DummyVecEnv([lambda: Monitor(env)]))` after a LOAD)."""
        from .compat import unwrap_env
        self.env = unwrap_env(env)
        self._last_obs = None
