"""The evaluation path (tester.py:41-63): an `Agent` plays the ego seat through
`set_ego_extractor(identity)` / `get_action(obs, False)` / `step`.  Reference numbers come from the reference's
own `run_test` executed on its own classes (tests/golden/make_golden_tester.py)."""
import os

import numpy as np

from pantheonrl_b200.common.agents import Agent
from pantheonrl_b200.envs.liar import LiarDefaultAgent
from test_wrappers_cpu import ReplayLiar


class Script(Agent):
    def __init__(self, actions):
        self.actions, self.k, self.seen = actions, 0, []

    def get_action(self, obs, record=True):
        self.seen.append((np.asarray(obs.obs).copy(), bool(record)))
        a = self.actions[self.k % len(self.actions)]
        self.k += 1
        return a

    def update(self, reward, done):
        pass


def test_run_test_loop_on_our_classes():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tester.npz"))
    env = ReplayLiar(g["resets"])
    env.add_partner_agent(LiarDefaultAgent())
    ego = Script(list(g["ego_script"]))
    env.set_ego_extractor(lambda obs: obs)  # the ego is an Agent: it gets Observation objects
    rewards = []
    for _ in range(25):
        obs, done, reward = env.reset(), False, 0
        env.render()
        while not done:
            obs, r, done, _ = env.step(ego.get_action(obs, False))
            reward += r
        rewards.append(reward)
    env.close()
    assert sum(rewards) / 25 == float(g["avg"]) and np.std(rewards) == float(g["std"])
    assert np.array_equal(np.array([o for o, _ in ego.seen]), g["ego_seen"])
    assert not any(r for _, r in ego.seen) and not g["ego_record_flags"].any()
