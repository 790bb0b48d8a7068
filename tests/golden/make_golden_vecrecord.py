"""Golden fixture: the reference's SimultaneousRecorder around its RPSEnv with scripted players
(authoring container only).  tests/test_vec_record_cpu.py cuts the same transitions out of rollout buffers."""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from pantheonrl.common.agents import Agent  # noqa: E402
from pantheonrl.common.wrappers import recorder_wrap  # noqa: E402
from pantheonrl.envs.rpsgym.rps import RPSEnv  # noqa: E402


class Script(Agent):
    def __init__(self, actions):
        self.actions, self.k = actions, 0

    def get_action(self, obs, record=True):
        a = self.actions[self.k % len(self.actions)]
        self.k += 1
        return a

    def update(self, reward, done):
        pass


def main():
    ego_s, alt_s, T = [0, 1, 2, 2, 1, 0, 0], [1, 1, 0, 2, 2], 33
    env = RPSEnv()
    env.add_partner_agent(Script(alt_s))
    env = rec = recorder_wrap(env)
    env.reset()
    for t in range(T):
        _, _, d, _ = env.step(ego_s[t % len(ego_s)])
        if d:
            env.reset()
    tr = rec.get_transitions()
    b = io.BytesIO()
    tr.write_transition(b)
    np.savez_compressed(os.path.join(HERE, "vec_record_rps.npz"), egoobs=tr.egoobs, egoacts=tr.egoacts, altobs=tr.altobs,
                        altacts=tr.altacts, flags=tr.flags, npy=np.frombuffer(b.getvalue(), np.uint8),
                        ego_script=np.array(ego_s), alt_script=np.array(alt_s), T=np.array(T))
    print(tr.egoobs.shape, tr.egoacts.shape, tr.flags[:5])


if __name__ == "__main__":
    main()
