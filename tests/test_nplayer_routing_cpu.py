"""MultiAgentEnv with THREE players, the ego in the middle seat, two candidates per partner slot and random
resampling: our host class against an event log of the reference's own class on the same game
(tests/golden/make_golden_nplayer.py; game logic shared through tests/golden/toy_games.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from toy_games import ThreePlayerLogic  # noqa: E402

from pantheonrl_b200.common.agents import Agent  # noqa: E402
from pantheonrl_b200.common.multiagentenv import MultiAgentEnv, PlayerException  # noqa: E402


from pantheonrl_b200.common.observation import Observation  # noqa: E402


class Game(ThreePlayerLogic, MultiAgentEnv):
    OBS = Observation

    def __init__(self, partners, **kw):
        MultiAgentEnv.__init__(self, ego_ind=1, n_players=3, partners=partners, **kw)


class Rec(Agent):
    def __init__(self, ident, log):
        self.ident, self.log, self.k = ident, log, 0

    def get_action(self, obs, record=True):
        a = (self.ident + self.k) % 3
        self.k += 1
        o = obs.obs
        self.log.append([0, self.ident, int(o[0]), int(o[1]), int(o[2]), a, 0.0, 0])
        return a

    def update(self, reward, done):
        self.log.append([1, self.ident, 0, 0, 0, 0, float(reward), int(done)])


def test_three_player_routing_matches_the_reference():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nplayer.npz"))
    log = []
    env = Game([[Rec(10, log), Rec(11, log)], [Rec(20, log), Rec(21, log)]])
    np.random.seed(9)
    ego = []
    for ep in range(12):
        o = env.reset()
        ego.append([2, ep, int(o[0]), int(o[1]), int(o[2]), 0, 0.0, 0] + list(env.partnerids))
        k = 0
        while True:
            a = (ep + k) % 3
            k += 1
            o, r, d, info = env.step(a)
            ego.append([3, a, int(o[0]), int(o[1]), int(o[2]), 0, float(r), int(d)] + list(info["_partnerid"]))
            if d:
                break
    assert np.array_equal(np.array(log, np.float64), g["partner_log"])   # every get_action / update of every partner
    assert np.array_equal(np.array(ego, np.float64), g["ego_log"])       # every obs / reward / done / partner ids the ego saw
    assert len({tuple(r[-2:]) for r in ego}) > 2                         # resampling really switched partners


def test_player_exceptions():
    with pytest.raises(PlayerException):
        Game([[Rec(1, [])]])                       # two partner slots needed
    with pytest.raises(PlayerException):
        Game([[Rec(1, [])], []])                   # empty slot
    with pytest.raises(PlayerException):
        Game([[Rec(1, [])], [Rec(2, [])]], resample_policy="robin")  # round robin is 2-player only
    env = Game([[Rec(1, [])], [Rec(2, [])]])
    with pytest.raises(PlayerException):
        env.add_partner_agent(Rec(3, []), player_num=1)               # seat 1 is the ego's
