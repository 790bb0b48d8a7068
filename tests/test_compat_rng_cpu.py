"""Reference-RNG mode, host side (no GPU): the who-starts coin and the dice drawn from np.random in the
reference's order reproduce what the reference's own LiarEnv / MultiAgentEnv drew under np.random.seed(10)
(tests/golden/compat_liar.npz, recorded by tests/golden/make_golden_compat.py from the reference's classes)."""
import os

import numpy as np

from pantheonrl_b200 import rng_mode
from pantheonrl_b200.envs.liar import LiarEnv, reference_hands

GOLD = os.path.join(os.path.dirname(__file__), "golden", "compat_liar.npz")


def test_coin_and_dice_stream_equals_the_reference():
    g = np.load(GOLD)
    np.random.seed(10)
    env = LiarEnv(rng="reference", device="cpu")  # draw_ego_first needs no device in this mode
    for k, want in enumerate(g["resets"]):
        ego_first = env.draw_ego_first()          # multiagentenv.py:325
        hands = reference_hands(np.random)        # liar.py:97-100
        assert int(ego_first) == want[0] and np.array_equal(hands, want[1:13]), k
    assert np.random.randint(1 << 30) == g["np_state_after"][0]  # not one draw more or less than the reference


def test_mode_switch():
    assert rng_mode.get_rng_mode() == "philox" and LiarEnv(device="cpu").rng == "philox"
    rng_mode.set_rng_mode("reference")
    try:
        assert LiarEnv(device="cpu").rng == "reference"
    finally:
        rng_mode.set_rng_mode("philox")
