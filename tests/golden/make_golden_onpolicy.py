"""Golden fixture for the plugin itself: the reference's OWN `OnPolicyAgent`
(pantheonrl/common/agents.py:82-208, executed verbatim) playing the partner's seat inside the
reference's own MultiAgentEnv (LiarEnv, RPSEnv), with a stand-in for the SB3 model that records
what the agent does to it (authoring container only: python tests/golden/make_golden_onpolicy.py).

Recorded per `model.train()` the agent triggers (lazily, inside the get_action AFTER its buffer
filled, agents.py:126): the rows it had written with `rollout_buffer.add` (obs, action,
episode_start, value, log_prob), the rewards it had accumulated into `buf.rewards[pos - 1]`
(agents.py:198), and the `compute_returns_and_advantage(last_values=, dones=)` arguments — the
bootstrap quirk of SURVEY.md Appendix B.1 — plus the ego-side trace and the dice, so that
tests/test_onpolicy_agent_cpu.py can replay the same scripts through OUR MultiAgentEnv /
OnPolicyAgent classes on a rule-equivalent CPU game.
"""
import os
import sys

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import pantheonrl.common.util as ref_util  # noqa: E402
ref_util.obs_as_tensor = lambda obs, device: th.as_tensor(obs)


class CaptureLogger:
    def __init__(self):
        self.dumps, self.kv = [], {}

    def record(self, key, value, exclude=None):
        self.kv[key] = value

    def dump(self, step=0):
        self.dumps.append((step, dict(self.kv)))
        self.kv = {}


sys.modules["stable_baselines3.common.utils"].configure_logger = lambda *a, **k: CaptureLogger()
sys.modules["stable_baselines3.common.utils"].safe_mean = lambda xs: float(np.mean(xs)) if len(xs) else float("nan")
import pantheonrl.common.agents as ref_agents  # noqa: E402
ref_agents.configure_logger = sys.modules["stable_baselines3.common.utils"].configure_logger
ref_agents.safe_mean = sys.modules["stable_baselines3.common.utils"].safe_mean
from pantheonrl.envs.liargym.liar import LiarEnv  # noqa: E402
from pantheonrl.envs.rpsgym.rps import RPSEnv  # noqa: E402


class FakeBuffer:
    def __init__(self, n_steps, obs_shape):
        self.n, self.obs_shape = n_steps, obs_shape
        self.rewards = np.zeros((n_steps, 1), np.float32)
        self.reset()
        self.gae_calls = []

    def reset(self):
        self.pos, self.rows = 0, []
        self.rewards[:] = 0

    def add(self, obs, actions, rewards, episode_starts, values, log_probs):
        self.rows.append((np.asarray(obs).reshape(-1).copy(), np.asarray(actions).reshape(-1).copy(),
                          float(episode_starts[0]), float(values.reshape(-1)[0]), float(log_probs.reshape(-1)[0])))
        self.rewards[self.pos] = np.asarray(rewards)
        self.pos += 1

    def compute_returns_and_advantage(self, last_values, dones):
        self.gae_calls.append((float(last_values.reshape(-1)[0]), float(dones)))


class ScriptedPolicy:
    """forward() returns the next scripted action with value = 0.25 * k and log_prob = -0.5 * k."""

    def __init__(self, obs_space, act_space, actions):
        self.observation_space, self.action_space, self.device = obs_space, act_space, "cpu"
        self.actions, self.k = actions, 0

    def forward(self, obs_tensor):
        a = np.asarray(self.actions[self.k % len(self.actions)]).reshape(1, -1)
        k = self.k
        self.k += 1
        if self.action_space.shape == ():
            a = a.reshape(1)
        return th.as_tensor(a), th.tensor([[0.25 * k]]), th.tensor([-0.5 * k])


class FakeModel:
    verbose, use_sde, sde_sample_freq = 0, False, -1

    def __init__(self, env, n_steps, actions):
        self.n_steps, self.action_space, self.observation_space = n_steps, env.action_space, env.observation_space
        self.policy = ScriptedPolicy(env.observation_space, env.action_space, actions)
        self.rollout_buffer = FakeBuffer(n_steps, env.observation_space.shape)
        self.trains, self.ep_info_buffer = [], None

    def set_logger(self, logger):
        self.logger = logger

    def train(self):
        b = self.rollout_buffer
        self.trains.append(dict(rows=list(b.rows), rewards=b.rewards[:b.pos, 0].copy(), gae=b.gae_calls[-1],
                                policy_calls=self.policy.k))


class RecordingLiar(LiarEnv):
    def __init__(self):
        super().__init__()
        self.resets = []

    def multi_reset(self, egofirst):
        o = super().multi_reset(egofirst)
        self.resets.append([int(egofirst)] + [int(x) for x in self.egohand] + [int(x) for x in self.althand])
        return o


def run(env, agent, model, ego_actions, n_steps):
    ego_obs, rews, dones = [], [], []
    o = env.reset()
    for t in range(n_steps):
        ego_obs.append(np.asarray(o).reshape(-1).copy())
        o, r, d, _ = env.step(ego_actions[t % len(ego_actions)])
        rews.append(r)
        dones.append(d)
        if d:
            o = env.reset()
    return np.array(ego_obs), np.array(rews, np.float64), np.array(dones)


def pack(prefix, out, model, agent, ego):
    T = model.trains
    out[prefix + "n_trains"] = np.array(len(T))
    out[prefix + "train_policy_calls"] = np.array([t["policy_calls"] for t in T])
    out[prefix + "gae_last_value"] = np.array([t["gae"][0] for t in T])
    out[prefix + "gae_dones"] = np.array([t["gae"][1] for t in T])
    out[prefix + "row_obs"] = np.array([[r[0] for r in t["rows"]] for t in T])
    out[prefix + "row_act"] = np.array([[r[1] for r in t["rows"]] for t in T])
    out[prefix + "row_start"] = np.array([[r[2] for r in t["rows"]] for t in T])
    out[prefix + "row_value"] = np.array([[r[3] for r in t["rows"]] for t in T])
    out[prefix + "row_logp"] = np.array([[r[4] for r in t["rows"]] for t in T])
    out[prefix + "row_reward"] = np.array([t["rewards"] for t in T])
    out[prefix + "agent_counters"] = np.array([agent.n_steps, agent.num_timesteps, agent.iteration])
    out[prefix + "pending_rows"] = np.array(len(model.rollout_buffer.rows))
    out[prefix + "log_steps"] = np.array([s for s, _ in model.logger.dumps])
    out[prefix + "log_ep_len"] = np.array([kv.get("rollout/ep_len_mean", np.nan) for _, kv in model.logger.dumps])
    out[prefix + "log_ep_rew"] = np.array([kv.get("rollout/ep_rew_mean", np.nan) for _, kv in model.logger.dumps])
    out[prefix + "ego_obs"], out[prefix + "ego_rew"], out[prefix + "ego_done"] = ego


def main():
    out = {}
    rng = np.random.RandomState(11)
    # ---- Liar's Dice: partner = OnPolicyAgent, n_steps = 5, many episodes (openings by the partner included)
    ego_script = [np.array([rng.randint(6), c]) for c in (1, 3, 5, 7, 9, 11, 2, 11, 4)] + [np.array([6, 11])]
    alt_script = [np.array([rng.randint(6), c]) for c in (2, 4, 6, 8, 10, 1, 11, 3)] + [np.array([6, 11])]
    np.random.seed(123)
    env = RecordingLiar()
    model = FakeModel(env, 5, alt_script)
    agent = ref_agents.OnPolicyAgent(model, log_interval=1)
    env.add_partner_agent(agent)
    ego = run(env, agent, model, ego_script, 80)
    pack("liar_", out, model, agent, ego)
    out["liar_resets"], out["liar_ego_script"], out["liar_alt_script"] = np.array(env.resets), np.array(ego_script), np.array(alt_script)
    # ---- RPS: every step ends an episode
    env = RPSEnv()
    model = FakeModel(env, 4, [0, 2, 1, 1, 2])
    agent = ref_agents.OnPolicyAgent(model, log_interval=2)
    env.add_partner_agent(agent)
    ego = run(env, agent, model, [1, 0, 2, 2, 1, 0, 0], 23)
    pack("rps_", out, model, agent, ego)
    out["rps_ego_script"], out["rps_alt_script"] = np.array([1, 0, 2, 2, 1, 0, 0]), np.array([0, 2, 1, 1, 2])
    np.savez_compressed(os.path.join(HERE, "onpolicy_agent.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
