"""Golden fixture for the small agents of the plugin surface, from the reference's own classes
(authoring container only): `LiarDefaultAgent` (liar.py:29-42), `RPSWeightedAgent` (rps.py:12-28),
`StaticPolicyAgent` (agents.py:54-79) with a scripted policy."""
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
from pantheonrl.common.agents import StaticPolicyAgent  # noqa: E402
from pantheonrl.common.observation import Observation  # noqa: E402
from pantheonrl.envs.liargym.liar import LiarDefaultAgent, LiarEnv  # noqa: E402
from pantheonrl.envs.rpsgym.rps import RPSWeightedAgent  # noqa: E402


def main():
    out = {}
    rng = np.random.RandomState(3)
    # LiarDefaultAgent on observations of real games (random play) and on random vectors of the space
    env, agent, obs_l = LiarEnv(), LiarDefaultAgent(), []
    np.random.seed(5)
    for _ in range(60):
        ef = bool(rng.rand() < 0.5)
        o = env.multi_reset(ef)
        turn = ef
        while True:
            obs_l.append(np.asarray(o).copy())
            o, _, d, _ = (env.ego_step if turn else env.alt_step)(np.array([rng.randint(7), rng.randint(12)]))
            turn = not turn
            if d:
                break
    for _ in range(200):
        obs_l.append(np.concatenate([rng.randint(0, 7, 6), np.stack([rng.randint(0, 7, 12), rng.randint(0, 12, 12)], 1).reshape(-1)]))
    out["liar_obs"] = np.array(obs_l)
    out["liar_act"] = np.array([agent.get_action(Observation(o)) for o in obs_l])
    # RPSWeightedAgent: seeded global np.random, several weightings
    w = [(1, 1, 1), (0, 0, 0), (3, 1, 0), (0, 5, 2), (1, 0, 0)]
    acts = []
    for i, (r, p, s) in enumerate(w):
        np.random.seed(40 + i)
        a = RPSWeightedAgent(r, p, s)
        acts.append([a.get_action(Observation(np.array([0]))) for _ in range(64)])
    out["rps_weights"], out["rps_act"] = np.array(w), np.array(acts)
    # StaticPolicyAgent: whatever policy.forward returns, first row, no recording, update ignored

    class Pol:
        device = "cpu"

        def __init__(self, space_shape, actions):
            import gym
            self.observation_space = gym.spaces.MultiDiscrete([7] * 30)
            self.action_space = gym.spaces.MultiDiscrete([7, 12])
            self.actions, self.k, self.seen = actions, 0, []

        def forward(self, obs_tensor):
            self.seen.append(obs_tensor.numpy().copy())
            a = self.actions[self.k % len(self.actions)]
            self.k += 1
            return th.as_tensor(np.asarray(a).reshape(1, -1)), th.zeros(1, 1), th.zeros(1)

    pol = Pol((30,), [np.array([1, 2]), np.array([6, 11]), np.array([0, 0])])
    sp = StaticPolicyAgent(pol)
    got = []
    for i in range(5):
        got.append(sp.get_action(Observation(out["liar_obs"][i]), record=bool(i % 2)))
        sp.update(1.0, bool(i % 2))
    out["static_act"] = np.array(got)
    out["static_seen_shape"] = np.array(pol.seen[0].shape)
    np.savez_compressed(os.path.join(HERE, "small_agents.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
