"""The ego's rollout loop of the facade (`pantheonrl_b200.ppo.collect_rollouts_single_env`, host code)
against a trace of the reference's in-tree copy of SB3's `collect_rollouts`
(pantheonrl/algos/adap/adap_learn.py:377-473, executed verbatim by tests/golden/make_golden_collect.py):
which observation each stored row carries, the episode_start latch, the rewards, the bootstrap
arguments, three rollouts in a row (state carried over), around OUR MultiAgentEnv host classes."""
import os
from collections import deque

import numpy as np

from pantheonrl_b200.common.agents import Agent
from pantheonrl_b200.ppo import collect_rollouts_single_env
from test_wrappers_cpu import ReplayLiar


class Script(Agent):
    def __init__(self, actions):
        self.actions, self.k = actions, 0

    def get_action(self, obs, record=True):
        a = self.actions[self.k % len(self.actions)]
        self.k += 1
        return a

    def update(self, reward, done):
        pass


class ScriptedPolicy:
    def __init__(self, actions):
        self.actions, self.k = actions, 0

    def forward(self, obs):
        a = np.asarray(self.actions[self.k % len(self.actions)]).reshape(1, -1)
        k = self.k
        self.k += 1
        return a, 0.25 * k, -0.5 * k

    def predict_values(self, obs):
        return self.forward(obs)[1]  # like DevicePolicy: one forward, one RNG tick


class RecBuffer:
    def __init__(self):
        self.rollouts = []

    def reset(self):
        self.rollouts.append(dict(rows=[], gae=None))

    def add(self, obs, action, reward, episode_start, value, log_prob):
        self.rollouts[-1]["rows"].append((np.asarray(obs).reshape(-1).copy(), np.asarray(action).reshape(-1).copy(),
                                          float(reward), float(episode_start), float(value), float(log_prob)))

    def compute_returns_and_advantage(self, last_values, dones):
        self.rollouts[-1]["gae"] = (float(last_values), float(dones))


class Algo:
    num_timesteps = 0


def test_collect_rollouts_matches_the_in_tree_copy():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "collect_rollouts.npz"))
    n_steps, n_rollouts = (int(x) for x in g["hp"])
    env = ReplayLiar(g["resets"])
    env.add_partner_agent(Script(list(g["alt_script"])))
    policy, buf, algo = ScriptedPolicy(list(g["ego_script"])), RecBuffer(), Algo()
    algo._last_obs, algo._last_start, algo._ep, algo.ep_info_buffer = env.reset(), True, [0.0, 0], deque(maxlen=100)
    for _ in range(n_rollouts):
        collect_rollouts_single_env(algo, env, policy, buf, n_steps)
    R = buf.rollouts
    assert len(R) == n_rollouts and algo.num_timesteps == int(g["num_timesteps"]) and policy.k == int(g["policy_calls"])
    for i, x in enumerate(R):
        assert np.array_equal(np.array([r[0] for r in x["rows"]]), g["row_obs"][i])      # the obs the action was chosen on
        assert np.array_equal(np.array([r[1] for r in x["rows"]]), g["row_act"][i])
        assert [r[2] for r in x["rows"]] == g["row_rew"][i].tolist()
        assert [r[3] for r in x["rows"]] == g["row_start"][i].tolist()                    # previous step's done; first row True
        assert np.allclose([r[4] for r in x["rows"]], g["row_value"][i]) and np.allclose([r[5] for r in x["rows"]], g["row_logp"][i])
        assert x["gae"][0] == g["gae_last_value"][i] and x["gae"][1] == g["gae_dones"][i]  # V(obs after the last step), its done
    eps = list(algo.ep_info_buffer)
    assert len(eps) == int(sum(g["row_start"].reshape(-1)[1:]) + g["gae_dones"][-1])       # one Monitor record per finished episode
    assert sum(e["l"] for e in eps) <= n_steps * n_rollouts


def test_adap_collect_rollouts_matches_the_reference():
    """`ADAP._collect` (pantheonrl_b200/adap.py, host code) against the reference's own `ADAP.collect_rollouts`
    (adap_learn.py:377-473) executed verbatim with context_size = 3: rows are observation ++ context, the context
    is resampled whenever an episode ends (after the auto-reset), the bootstrap value comes from one more
    policy.forward."""
    import torch
    from pantheonrl_b200 import adap

    class ContextPolicy(ScriptedPolicy):
        def __init__(self, actions):
            super().__init__(actions)
            self.context, self.sets = torch.tensor([[0.125, 0.25, 0.5]]), []

        def get_context(self):
            return self.context

        def set_context(self, c):
            self.context = c
            self.sets.append(self.k)

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "collect_rollouts_adap.npz"))
    n_steps, n_rollouts = (int(x) for x in g["hp"])
    env = ReplayLiar(g["resets"])
    env.add_partner_agent(Script(list(g["alt_script"])))
    policy, buf = ContextPolicy(list(g["ego_script"])), RecBuffer()
    drawn = []

    class Algo_:
        num_timesteps = 0

        def _sample_context(self):  # the golden run's counting sampler: k-th context = [k, k + 0.5, -k]
            k = len(drawn) + 1
            drawn.append(k)
            return torch.tensor([[float(k), k + 0.5, -float(k)]])
    algo = Algo_()
    algo.policy, algo.n_steps = policy, n_steps
    algo._last_obs, algo._last_start, algo._ep, algo.ep_info_buffer = env.reset(), True, [0.0, 0], deque(maxlen=100)
    for _ in range(n_rollouts):
        adap.ADAP._collect(algo, env, buf)
    R = buf.rollouts
    assert len(R) == n_rollouts and algo.num_timesteps == int(g["num_timesteps"]) and policy.k == int(g["policy_calls"])
    assert policy.sets == g["context_sets"].tolist() and len(drawn) == int(g["n_drawn"])
    assert tuple(g["full_obs_shape"]) == (33,)
    for i, x in enumerate(R):
        assert np.array_equal(np.array([r[0] for r in x["rows"]]), g["row_obs"][i])  # observation ++ the context it was chosen under
        assert np.array_equal(np.array([r[1] for r in x["rows"]]), g["row_act"][i])
        assert [r[2] for r in x["rows"]] == g["row_rew"][i].tolist()
        assert [r[3] for r in x["rows"]] == g["row_start"][i].tolist()
        assert np.allclose([r[4] for r in x["rows"]], g["row_value"][i]) and np.allclose([r[5] for r in x["rows"]], g["row_logp"][i])
        assert x["gae"][0] == g["gae_last_value"][i] and x["gae"][1] == g["gae_dones"][i]


def test_modular_collect_rollouts_matches_the_reference():
    """`ModularAlgorithm._collect_partner` (pantheonrl_b200/modular.py, host code) against the reference's own
    `ModularAlgorithm.collect_rollouts` (modular/learn.py:155-218) executed verbatim for two partners x two
    iterations: the rows, which partner answered (set_partnerid before every step, the env's own round robin at
    every reset), the bootstrap from the LAST forward's value.  One documented difference: the reference adds the
    first row of every rollout with episode_start = None (`self._last_dones = None`, learn.py:181) — ours carries
    the previous step's done flag."""
    from pantheonrl_b200 import modular

    class PartnerPolicy(ScriptedPolicy):
        def __init__(self, actions):
            super().__init__(actions)
            self.partner_of_call = []

        def forward(self, obs, partner_idx=0):
            self.partner_of_call.append(int(partner_idx))
            return super().forward(obs)

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "collect_rollouts_modular.npz"))
    n_steps, n_iter = (int(x) for x in g["hp"])
    env = ReplayLiar(g["resets"])
    partners = [Script(list(g["alt_script0"])), Script(list(g["alt_script1"]))]
    for p in partners:
        env.add_partner_agent(p)
    policy, bufs, algo = PartnerPolicy(list(g["ego_script"])), [RecBuffer(), RecBuffer()], Algo()
    algo.policy, algo.n_steps = policy, n_steps
    algo._last_obs, algo._last_start, algo._ep, algo.ep_info_buffer = env.reset(), True, [0.0, 0], deque(maxlen=100)
    calls = []
    for _ in range(n_iter):
        for q in range(2):
            env.set_partnerid(q)
            modular.ModularAlgorithm._collect_partner(algo, env, bufs[q], q)
            calls.append([p.k for p in partners])
    assert algo.num_timesteps == int(g["num_timesteps"]) and policy.k == int(g["policy_calls"])
    assert policy.partner_of_call == g["partner_of_call"].tolist() and calls == g["partner_calls"].tolist()
    prev_done = True
    for it in range(n_iter):
        for q in range(2):
            x = bufs[q].rollouts[it]
            assert np.array_equal(np.array([r[0] for r in x["rows"]]), g[f"p{q}_row_obs"][it])
            assert np.array_equal(np.array([r[1] for r in x["rows"]]), g[f"p{q}_row_act"][it])
            assert [r[2] for r in x["rows"]] == g[f"p{q}_row_rew"][it].tolist()
            starts, want = [r[3] for r in x["rows"]], g[f"p{q}_row_start"][it].tolist()
            assert want[0] == -1.0 and starts[0] == float(prev_done)   # reference: None; ours: the previous step's done
            assert starts[1:] == want[1:]
            assert np.allclose([r[4] for r in x["rows"]], g[f"p{q}_row_value"][it])
            assert x["gae"][0] == g[f"p{q}_gae_last_value"][it] and x["gae"][1] == g[f"p{q}_gae_dones"][it]
            prev_done = bool(x["gae"][1])
