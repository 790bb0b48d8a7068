"""The small agents of the plugin surface against the reference's own classes
(tests/golden/make_golden_agents.py): LiarDefaultAgent, RPSWeightedAgent, StaticPolicyAgent."""
import os

import numpy as np

from pantheonrl_b200.common.agents import StaticPolicyAgent
from pantheonrl_b200.common.observation import Observation
from pantheonrl_b200.envs.liar import LiarDefaultAgent
from pantheonrl_b200.envs.rps import RPSWeightedAgent

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_agents.npz"))


def test_liar_default_agent():
    agent = LiarDefaultAgent()
    got = np.array([agent.get_action(Observation(o)) for o in G["liar_obs"]])
    assert np.array_equal(got, G["liar_act"])
    assert (got == [6, 11]).all(axis=1).any() and not (got == [6, 11]).all(axis=1).all()  # both branches were hit


def test_rps_weighted_agent_consumes_the_same_global_stream():
    for i, (r, p, s) in enumerate(G["rps_weights"]):
        np.random.seed(40 + i)
        a = RPSWeightedAgent(int(r), int(p), int(s))
        assert [a.get_action(Observation(np.array([0]))) for _ in range(64)] == G["rps_act"][i].tolist()


def test_static_policy_agent():
    class Pol:
        def __init__(self, actions):
            self.actions, self.k, self.seen = actions, 0, []

        def forward(self, obs):
            self.seen.append(np.asarray(obs).copy())
            a = self.actions[self.k % len(self.actions)]
            self.k += 1
            return np.asarray(a).reshape(1, -1), 0.0, 0.0

    pol = Pol([np.array([1, 2]), np.array([6, 11]), np.array([0, 0])])
    sp = StaticPolicyAgent(pol)
    got = []
    for i in range(5):
        got.append(sp.get_action(Observation(G["liar_obs"][i]), record=bool(i % 2)))
        sp.update(1.0, bool(i % 2))
    assert np.array_equal(np.array(got), G["static_act"]) and pol.k == 5
