"""Golden fixture for the evaluation path: the reference's own `run_test` loop
(tester.py:41-63, exec'd from the file's source lines) on the reference's LiarEnv with the ego played
by an `Agent` (identity ego extractor: the ego receives Observation objects) against a scripted partner.
Authoring container only:  python tests/golden/make_golden_tester.py"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from pantheonrl.common.agents import Agent  # noqa: E402
from pantheonrl.envs.liargym.liar import LiarDefaultAgent, LiarEnv  # noqa: E402


class Script(Agent):
    def __init__(self, actions):
        self.actions, self.k, self.seen = actions, 0, []

    def get_action(self, obs, record=True):
        assert record is False or record is True
        self.seen.append((np.asarray(obs.obs).copy(), bool(record)))
        a = self.actions[self.k % len(self.actions)]
        self.k += 1
        return a

    def update(self, reward, done):
        pass


class RecordingLiar(LiarEnv):
    def __init__(self):
        super().__init__()
        self.resets, self.closed = [], 0

    def multi_reset(self, egofirst):
        o = super().multi_reset(egofirst)
        self.resets.append([int(egofirst)] + [int(x) for x in self.egohand] + [int(x) for x in self.althand])
        return o

    def close(self):
        self.closed += 1


def main():
    src = open(os.path.join(ref_shim.REF, "tester.py")).read().splitlines()
    start = next(i for i, ln in enumerate(src) if ln.startswith("def run_test("))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("if __name__") or src[i].startswith("def "))
    ns = {"np": np, "sleep": lambda s: None}
    exec("\n".join(src[start:end]), ns)  # the reference's own function
    rng = np.random.RandomState(2)
    ego_script = [np.array([rng.randint(6), c]) for c in (1, 2, 4, 6, 8, 10, 11)] + [np.array([6, 11])]
    np.random.seed(17)
    env = RecordingLiar()
    env.add_partner_agent(LiarDefaultAgent())
    ego = Script(ego_script)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ns["run_test"](ego, env, 25, False)
    printed = buf.getvalue().strip().splitlines()
    np.savez_compressed(os.path.join(HERE, "tester.npz"), resets=np.array(env.resets), ego_script=np.array(ego_script),
                        ego_seen=np.array([o for o, _ in ego.seen]), ego_record_flags=np.array([r for _, r in ego.seen]),
                        closed=np.array(env.closed), avg=np.array(float(printed[0].split(":")[1])),
                        std=np.array(float(printed[1].split(":")[1])))
    print(printed, len(ego.seen))


if __name__ == "__main__":
    main()
