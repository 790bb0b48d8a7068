"""Golden fixture for the ego's rollout loop: the reference's in-tree copy of SB3's
`OnPolicyAlgorithm.collect_rollouts` (pantheonrl/algos/adap/adap_learn.py:377-473, `ADAP.collect_rollouts`),
executed verbatim and unbound on a duck-typed `self` with context size 0, around the reference's own
MultiAgentEnv (LiarEnv with a scripted partner) behind a stand-in for SB3's DummyVecEnv (SB3 is not
installable here; the stand-in restates SURVEY.md Appendix A7: auto-reset on done, leading env axis).
Authoring container only:  python tests/golden/make_golden_collect.py

Records, for three consecutive rollouts of 7 steps: every `rollout_buffer.add(obs, actions, rewards,
episode_starts, values, log_probs)` and the `compute_returns_and_advantage(last_values=, dones=)`
arguments.  tests/test_collect_rollouts_cpu.py replays the scripts through
pantheonrl_b200.ppo.collect_rollouts_single_env.
"""
import os
import sys
import types

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()


def stub(name, **attrs):
    m = sys.modules.get(name) or types.ModuleType(name)
    sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


stub("stable_baselines3.common.type_aliases", GymEnv=object, MaybeCallback=object, Schedule=object)
stub("stable_baselines3.common.utils", explained_variance=None, get_schedule_fn=None,
     obs_as_tensor=lambda obs, device: th.as_tensor(obs))
stub("stable_baselines3.common.vec_env", VecEnv=object)
stub("stable_baselines3.common.callbacks", BaseCallback=object)
stub("stable_baselines3.common.buffers", RolloutBuffer=object)
_drawn = []


def counting_sampler(ctx_size, num, torch):
    """Deterministic stand-in for SAMPLERS[...]: the k-th sampled context is [k + 1, k + 1.5, -(k + 1)]."""
    k = len(_drawn) + 1
    _drawn.append(k)
    return th.tensor([[float(k), k + 0.5, -float(k)]])[:, :ctx_size]


stub("pantheonrl.algos.adap.util", SAMPLERS={"none": lambda ctx_size, num, torch: None, "count": counting_sampler},
     get_context_kl_loss=None)
stub("pantheonrl.algos.adap.policies", AdapPolicy=object)
from pantheonrl.algos.adap import adap_learn  # noqa: E402
from pantheonrl.common.agents import Agent  # noqa: E402
from pantheonrl.envs.liargym.liar import LiarEnv  # noqa: E402


class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Script(Agent):
    def __init__(self, actions):
        self.actions, self.k = actions, 0

    def get_action(self, obs, record=True):
        a = self.actions[self.k % len(self.actions)]
        self.k += 1
        return a

    def update(self, reward, done):
        pass


class RecordingLiar(LiarEnv):
    def __init__(self):
        super().__init__()
        self.resets = []

    def multi_reset(self, egofirst):
        o = super().multi_reset(egofirst)
        self.resets.append([int(egofirst)] + [int(x) for x in self.egohand] + [int(x) for x in self.althand])
        return o


class OneEnvVec:
    """DummyVecEnv([lambda: env]) as SB3 1.7.0 defines it."""
    num_envs = 1

    def __init__(self, env):
        self.env = env

    def reset(self):
        return np.asarray(self.env.reset())[None]

    def step(self, actions):
        obs, rew, done, info = self.env.step(actions[0])
        if done:
            info = dict(info, terminal_observation=obs)
            obs = self.env.reset()
        return np.asarray(obs)[None], np.array([rew], np.float32), np.array([done]), [info]


class ScriptedPolicy:
    def __init__(self, actions):
        self.actions, self.k = actions, 0

    def forward(self, obs_tensor):
        a = np.asarray(self.actions[self.k % len(self.actions)]).reshape(1, -1)
        k = self.k
        self.k += 1
        return th.as_tensor(a), th.tensor([[0.25 * k]]), th.tensor([-0.5 * k])

    def get_context(self):
        return np.zeros(0)

    def set_context(self, ctx):
        pass


class RecBuffer:
    def __init__(self, obs_shape):
        self.obs_shape, self.rollouts = obs_shape, []

    def reset(self):
        self.rollouts.append(dict(rows=[], gae=None))

    def add(self, obs, actions, rewards, episode_starts, values, log_probs):
        self.rollouts[-1]["rows"].append((np.asarray(obs).reshape(-1).copy(), np.asarray(actions).reshape(-1).copy(),
                                          float(rewards[0]), float(np.asarray(episode_starts).reshape(-1)[0]),
                                          float(values.reshape(-1)[0]), float(log_probs.reshape(-1)[0])))

    def compute_returns_and_advantage(self, last_values, dones):
        self.rollouts[-1]["gae"] = (float(last_values.reshape(-1)[0]), float(np.asarray(dones).reshape(-1)[0]))


def main():
    import gym
    rng = np.random.RandomState(21)
    ego_script = [np.array([rng.randint(6), c]) for c in (1, 3, 5, 7, 9, 11, 2, 11, 4)] + [np.array([6, 11])]
    alt_script = [np.array([rng.randint(6), c]) for c in (2, 4, 6, 8, 10, 1, 11, 3)] + [np.array([6, 11])]
    np.random.seed(321)
    base = RecordingLiar()
    base.add_partner_agent(Script(alt_script))
    venv = OneEnvVec(base)
    buf = RecBuffer(base.observation_space.shape)
    cb = Data(on_rollout_start=lambda: None, on_rollout_end=lambda: None, update_locals=lambda l: None,
              on_step=lambda: True)
    algo = Data(_last_obs=venv.reset(), _last_episode_starts=np.ones((1,), dtype=bool), full_obs_shape=None,
                context_size=0, use_sde=False, sde_sample_freq=-1, policy=ScriptedPolicy(ego_script), device="cpu",
                action_space=gym.spaces.MultiDiscrete([7, 12]), num_timesteps=0, _update_info_buffer=lambda infos: None,
                context_sampler="none")
    n_steps, n_rollouts = 7, 3
    for _ in range(n_rollouts):
        assert adap_learn.ADAP.collect_rollouts(algo, venv, cb, buf, n_steps) is True  # <- the reference's own code
    R = buf.rollouts
    out = dict(
        row_obs=np.array([[r[0] for r in x["rows"]] for x in R]), row_act=np.array([[r[1] for r in x["rows"]] for x in R]),
        row_rew=np.array([[r[2] for r in x["rows"]] for x in R]), row_start=np.array([[r[3] for r in x["rows"]] for x in R]),
        row_value=np.array([[r[4] for r in x["rows"]] for x in R]), row_logp=np.array([[r[5] for r in x["rows"]] for x in R]),
        gae_last_value=np.array([x["gae"][0] for x in R]), gae_dones=np.array([x["gae"][1] for x in R]),
        num_timesteps=np.array(algo.num_timesteps), policy_calls=np.array(algo.policy.k), resets=np.array(base.resets),
        ego_script=np.array(ego_script), alt_script=np.array(alt_script), hp=np.array([n_steps, n_rollouts]))
    np.savez_compressed(os.path.join(HERE, "collect_rollouts.npz"), **out)
    print({k: v.shape for k, v in out.items()}, "starts:", out["row_start"].tolist(), "gae:", out["gae_last_value"], out["gae_dones"])


class ContextPolicy(ScriptedPolicy):
    """ScriptedPolicy that carries a context like AdapPolicy (set_context / get_context)."""

    def __init__(self, actions):
        super().__init__(actions)
        self.context, self.sets = th.tensor([[0.125, 0.25, 0.5]]), []

    def get_context(self):
        return self.context

    def set_context(self, ctx):
        self.context = ctx
        self.sets.append(self.k)  # how many forwards had happened when the context changed


def main_adap():
    """The same loop with context_size = 3: rows are observation ++ context, the context is resampled when an
    episode ends (adap_learn.py:441-455), the bootstrap value comes from policy.forward (:457-460)."""
    import gym
    rng = np.random.RandomState(22)
    ego_script = [np.array([rng.randint(6), c]) for c in (1, 3, 5, 7, 9, 11, 2, 11, 4)] + [np.array([6, 11])]
    alt_script = [np.array([rng.randint(6), c]) for c in (2, 4, 6, 8, 10, 1, 11, 3)] + [np.array([6, 11])]
    np.random.seed(322)
    base = RecordingLiar()
    base.add_partner_agent(Script(alt_script))
    venv = OneEnvVec(base)
    buf = RecBuffer(base.observation_space.shape)
    cb = Data(on_rollout_start=lambda: None, on_rollout_end=lambda: None, update_locals=lambda l: None,
              on_step=lambda: True)
    pol = ContextPolicy(ego_script)
    algo = Data(_last_obs=venv.reset(), _last_episode_starts=np.ones((1,), dtype=bool), full_obs_shape=None,
                context_size=3, use_sde=False, sde_sample_freq=-1, policy=pol, device="cpu",
                action_space=gym.spaces.MultiDiscrete([7, 12]), num_timesteps=0, _update_info_buffer=lambda infos: None,
                context_sampler="count")
    n_steps, n_rollouts = 9, 3
    for _ in range(n_rollouts):
        assert adap_learn.ADAP.collect_rollouts(algo, venv, cb, buf, n_steps) is True  # <- the reference's own code
    R = buf.rollouts
    out = dict(
        row_obs=np.array([[r[0] for r in x["rows"]] for x in R]), row_act=np.array([[r[1] for r in x["rows"]] for x in R]),
        row_rew=np.array([[r[2] for r in x["rows"]] for x in R]), row_start=np.array([[r[3] for r in x["rows"]] for x in R]),
        row_value=np.array([[r[4] for r in x["rows"]] for x in R]), row_logp=np.array([[r[5] for r in x["rows"]] for x in R]),
        gae_last_value=np.array([x["gae"][0] for x in R]), gae_dones=np.array([x["gae"][1] for x in R]),
        num_timesteps=np.array(algo.num_timesteps), policy_calls=np.array(pol.k), resets=np.array(base.resets),
        context_sets=np.array(pol.sets), n_drawn=np.array(len(_drawn)), full_obs_shape=np.array(algo.full_obs_shape),
        ego_script=np.array(ego_script), alt_script=np.array(alt_script), hp=np.array([n_steps, n_rollouts]))
    np.savez_compressed(os.path.join(HERE, "collect_rollouts_adap.npz"), **out)
    print("adap:", out["row_obs"].shape, "context changes after forwards", out["context_sets"].tolist(), "gae:",
          out["gae_last_value"], out["gae_dones"])


if __name__ == "__main__":
    main()
    main_adap()
