"""Rule-only game logic shared by a golden generator (mixed into the REFERENCE's MultiAgentEnv) and the test
that replays it (mixed into OURS): a 3-player game in which different subsets of players move each step."""
import numpy as np

SCHEDULE = [(0,), (1, 2), (0, 1, 2), (2,), (1,), (0, 2), (1,)]  # who moves at step t (player 1 is the ego)


class ThreePlayerLogic:
    """n_reset / n_step of a MultiAgentEnv with n_players = 3; rewards depend on the actions and the step."""
    LENGTH = len(SCHEDULE) - 1
    OBS = None  # the Observation class of whichever package the logic is mixed into

    def n_reset(self):
        self.t = 0
        players = SCHEDULE[0]
        return players, tuple(self.OBS(np.array([0, p, 9])) for p in players)

    def n_step(self, actions):
        movers = SCHEDULE[self.t]
        acts = [int(a) for a in actions]
        assert len(acts) == len(movers)
        self.t += 1
        rews = [0.0, 0.0, 0.0]
        for p, a in zip(movers, acts):
            rews[p] += a + 0.5 * self.t
            rews[(p + 1) % 3] -= 0.25 * a
        done = self.t >= self.LENGTH
        players = SCHEDULE[self.t]
        obs = tuple(self.OBS(np.array([self.t, p, sum(acts)])) for p in players)
        return players, obs, tuple(rews), done, {}
