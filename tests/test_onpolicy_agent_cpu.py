"""The plugin itself on the CPU: OUR `OnPolicyAgent` + `MultiAgentEnv` host classes against a trace of
the reference's OWN `OnPolicyAgent` (pantheonrl/common/agents.py:82-208, executed verbatim inside the
reference's MultiAgentEnv by tests/golden/make_golden_onpolicy.py) with a recording stand-in for the SB3
model: every row the agent writes, every reward it accumulates, WHEN it trains (lazily, inside the next
get_action), what it bootstraps from (Appendix B.1), what it logs, and what the ego sees."""
import os

import numpy as np
import pytest

import oracle
from pantheonrl_b200.common.agents import OnPolicyAgent
from pantheonrl_b200.common.multiagentenv import SimultaneousEnv
from pantheonrl_b200.spaces import Discrete
from test_wrappers_cpu import ReplayLiar


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "onpolicy_agent.npz"))


class CaptureLogger:
    output_formats = ["capture"]

    def __init__(self):
        self.dumps, self.kv = [], {}

    def record(self, key, value, exclude=None):
        self.kv[key] = value

    def dump(self, step=0):
        self.dumps.append((step, dict(self.kv)))
        self.kv = {}


class RecBuffer:
    """The facade's buffer surface (HostStagedBuffer): add / add_reward / reset / compute_returns_and_advantage."""

    def __init__(self):
        self.gae_calls = []
        self.reset()

    def reset(self):
        self.pos, self.rows, self.rewards = 0, [], []

    def add(self, obs, action, reward, episode_start, value, log_prob):
        self.rows.append((np.asarray(obs).reshape(-1).copy(), np.asarray(action).reshape(-1).copy(),
                          float(episode_start), float(value), float(log_prob)))
        self.rewards.append(float(reward))
        self.pos += 1

    def add_reward(self, reward):
        if self.pos > 0:
            self.rewards[self.pos - 1] += float(reward)

    def compute_returns_and_advantage(self, last_values, dones):
        self.gae_calls.append((float(last_values), float(dones)))


class ScriptedPolicy:
    def __init__(self, actions, discrete):
        self.actions, self.k, self.discrete = actions, 0, discrete

    def forward(self, obs):
        a = np.asarray(self.actions[self.k % len(self.actions)]).reshape(1, -1)
        k = self.k
        self.k += 1
        return (a.reshape(1) if self.discrete else a), 0.25 * k, -0.5 * k


class RecModel:
    verbose = 0

    def __init__(self, n_steps, actions, discrete):
        self.n_steps, self.policy, self.rollout_buffer = n_steps, ScriptedPolicy(actions, discrete), RecBuffer()
        self.trains, self.ep_info_buffer = [], None

    def set_logger(self, logger):
        self.logger = CaptureLogger()

    def train(self):
        b = self.rollout_buffer
        self.trains.append(dict(rows=list(b.rows), rewards=np.array(b.rewards), gae=b.gae_calls[-1],
                                policy_calls=self.policy.k))


class CpuRPS(SimultaneousEnv):
    def __init__(self):
        super().__init__()
        self.observation_space, self.action_space = Discrete(1), Discrete(3)

    def multi_reset(self):
        return np.array([0]), np.array([0])

    def multi_step(self, ego_action, alt_action):
        re, ra = oracle.rps_step([int(ego_action)], [int(alt_action)])
        return (np.array([0]), np.array([0])), (float(re[0]), float(ra[0])), True, {}


def run(env, ego_actions, n_steps):
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


@pytest.mark.parametrize("name", ["liar", "rps"])
def test_onpolicy_agent_matches_the_references_class(g, name):
    pre = name + "_"
    if name == "liar":
        env, n_steps, log_interval, T = ReplayLiar(g["liar_resets"]), 5, 1, 80
    else:
        env, n_steps, log_interval, T = CpuRPS(), 4, 2, 23
    model = RecModel(n_steps, list(g[pre + "alt_script"]), discrete=name == "rps")
    agent = OnPolicyAgent(model, log_interval=log_interval)
    env.add_partner_agent(agent)
    ego_obs, ego_rew, ego_done = run(env, list(g[pre + "ego_script"]), T)
    # what the ego saw
    assert np.array_equal(ego_obs, g[pre + "ego_obs"]) and np.array_equal(ego_rew, g[pre + "ego_rew"])
    assert np.array_equal(ego_done, g[pre + "ego_done"])
    # when the partner trained and what it bootstrapped from
    tr = model.trains
    assert len(tr) == int(g[pre + "n_trains"]) > 2
    assert [t["policy_calls"] for t in tr] == g[pre + "train_policy_calls"].tolist()   # lazily, inside the NEXT get_action
    assert np.allclose([t["gae"][0] for t in tr], g[pre + "gae_last_value"])          # value of the LAST STORED obs (B.1)
    assert [t["gae"][1] for t in tr] == g[pre + "gae_dones"].tolist()
    # every row it wrote, with the rewards it accumulated onto it
    for i, t in enumerate(tr):
        assert np.array_equal(np.array([r[0] for r in t["rows"]]), g[pre + "row_obs"][i])
        assert np.array_equal(np.array([r[1] for r in t["rows"]]), g[pre + "row_act"][i])
        assert [r[2] for r in t["rows"]] == g[pre + "row_start"][i].tolist()
        assert np.allclose([r[3] for r in t["rows"]], g[pre + "row_value"][i])
        assert np.allclose([r[4] for r in t["rows"]], g[pre + "row_logp"][i])
        assert np.array_equal(t["rewards"], g[pre + "row_reward"][i]), i
    assert [agent.n_steps, agent.num_timesteps, agent.iteration] == g[pre + "agent_counters"].tolist()
    assert len(model.rollout_buffer.rows) == int(g[pre + "pending_rows"])
    # what it logged (agents.py:132-153)
    assert [s for s, _ in model.logger.dumps] == g[pre + "log_steps"].tolist()
    assert np.allclose([kv["rollout/ep_len_mean"] for _, kv in model.logger.dumps], g[pre + "log_ep_len"], equal_nan=True)
    assert np.allclose([kv["rollout/ep_rew_mean"] for _, kv in model.logger.dumps], g[pre + "log_ep_rew"], equal_nan=True)
