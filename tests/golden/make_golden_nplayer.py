"""Golden fixture: the reference's OWN MultiAgentEnv (multiagentenv.py:21-285) driving a 3-player game
(tests/golden/toy_games.py) with the ego in the middle seat, two candidate partners per slot and the default
(random) resampling policy.  Authoring container only:  python tests/golden/make_golden_nplayer.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from pantheonrl.common.agents import Agent  # noqa: E402
from pantheonrl.common.multiagentenv import MultiAgentEnv  # noqa: E402
from toy_games import ThreePlayerLogic  # noqa: E402


from pantheonrl.common.observation import Observation  # noqa: E402


class Game(ThreePlayerLogic, MultiAgentEnv):
    OBS = Observation

    def __init__(self, partners):
        MultiAgentEnv.__init__(self, ego_ind=1, n_players=3, partners=partners)


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


def main():
    log = []
    partners = [[Rec(10, log), Rec(11, log)], [Rec(20, log), Rec(21, log)]]
    env = Game(partners)
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
    np.savez_compressed(os.path.join(HERE, "nplayer.npz"), partner_log=np.array(log, np.float64),
                        ego_log=np.array(ego, np.float64))
    print(len(log), len(ego))


if __name__ == "__main__":
    main()
