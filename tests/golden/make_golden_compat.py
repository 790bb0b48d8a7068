"""Golden fixture for the reference-RNG mode (pantheonrl_b200/rng_mode.py): the dice, coins, observations
and rewards the REFERENCE's own LiarEnv + MultiAgentEnv produce under np.random.seed(10) when the ego and
the partner play scripted actions (their own RandomStates, so that np.random is consumed by the env alone,
exactly as in `trainer.py LiarsDice-v0 ...` where the agents sample from torch's generator).

Run in the authoring container only:  python tests/golden/make_golden_compat.py  -> compat_liar.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from pantheonrl.common.agents import Agent  # noqa: E402
from pantheonrl.envs.liargym.liar import LiarEnv  # noqa: E402

T = 2048


class Scripted(Agent):
    def __init__(self, acts):
        self.acts, self.k = acts, 0

    def get_action(self, obs, record=True):
        a = self.acts[self.k]
        self.k += 1
        return a

    def update(self, reward, done):
        pass


def main():
    rs = np.random.RandomState(123)
    ego_acts = np.stack([rs.randint(0, 7, 3 * T), rs.randint(0, 12, 3 * T)], 1)
    alt_acts = np.stack([rs.randint(0, 7, 3 * T), rs.randint(0, 12, 3 * T)], 1)
    np.random.seed(10)
    env = LiarEnv()
    env.add_partner_agent(Scripted(alt_acts))
    resets, obs_l, rew_l, done_l = [], [], [], []

    # the who-starts coin is read right after n_reset: wrap it
    orig = env.n_reset

    def n_reset():
        out = orig()
        resets.append([int(out[0][0] == 0)] + list(env.egohand) + list(env.althand))
        return out
    env.n_reset = n_reset
    obs = env.reset()
    for t in range(T):
        obs_l.append(np.asarray(obs).copy())
        obs, r, d, _ = env.step(ego_acts[t])
        rew_l.append(r)
        done_l.append(d)
        if d:
            obs = env.reset()
    np.savez_compressed(os.path.join(HERE, "compat_liar.npz"), ego_acts=ego_acts[:T].astype(np.uint8),
                        alt_acts=alt_acts.astype(np.uint8), resets=np.array(resets, np.uint8),
                        obs=np.array(obs_l, np.uint8), rew=np.array(rew_l, np.float32),
                        done=np.array(done_l, np.uint8), final_obs=np.asarray(obs, np.uint8),
                        np_state_after=np.array([np.random.randint(1 << 30)]))
    print("resets", len(resets), "episodes", int(np.sum(done_l)))


if __name__ == "__main__":
    main()
