"""CPU tests: pin the oracle's rollout driver (who moves when, reward routing,
first-move hand-off, auto-reset with the partner's opening move, episode_start
latches) against event traces recorded from the reference's own
MultiAgentEnv / TurnBasedEnv / SimultaneousEnv classes (tests/golden)."""
import os

import numpy as np
import pytest

import oracle
from oracle import rollout as orc


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _pad4(a):
    out = np.zeros((a.shape[0], 4), np.uint8)
    out[:, :a.shape[1]] = a
    return out


@pytest.mark.parametrize("env_kind,fname,space_kw", [
    ("liar", "routing_liar.npz", oracle.LIAR_SPACE),
    ("rps", "routing_rps.npz", oracle.RPS_SPACE),
])
def test_scripted_rollout_reproduces_reference_trace(golden_dir, env_kind, fname, space_kw):
    g = _load(golden_dir, fname)
    T = g["ego_act"].shape[0]
    rows, latch = orc.partner_rows_from_events(g)
    alt_act = _pad4(np.array([r["act"] for r in rows], np.uint8))
    space = oracle.make_space(**space_kw)
    ego, alt, carry = orc.rollout(
        env_kind, space, None, None, N=1, T=T, script_ego_act=_pad4(g["ego_act"]),
        script_alt_act=alt_act, script_reset=g["reset_info"] if env_kind == "liar" else None,
        alt=orc.new_buffer(len(rows) + 4, 1, True))
    # ---- what the ego (SB3 collect_rollouts) saw
    nobs = g["ego_obs"].shape[1]
    assert np.array_equal(ego["obs"][:, 0, :nobs], g["ego_obs"])
    assert np.array_equal(ego["actions"][:, 0, :g["ego_act"].shape[1]], g["ego_act"])
    assert np.array_equal(ego["rewards"][:, 0], g["ego_rew"])
    starts = np.concatenate([[1.0], g["ego_done"][:-1].astype(np.float32)])
    assert np.array_equal(ego["episode_starts"][:, 0], starts)
    assert carry["ego_last_done"][0] == float(g["ego_done"][-1])
    # ---- what the partner (OnPolicyAgent) stored
    K = len(rows)
    # a turn-based trace that ends while the partner waits for the ego's next move leaves its last row
    # open: not counted, flagged, kept at row count[n] for the next rollout
    is_open = int(carry["flags"][0] >> 2) & 1
    assert alt["count"][0] + is_open == K
    if env_kind == "rps":
        assert not is_open
    assert np.array_equal(alt["obs"][:K, 0, :nobs], np.array([r["obs"] for r in rows]))
    assert np.array_equal(alt["rewards"][:K, 0], np.array([r["rew"] for r in rows], np.float32))
    assert np.array_equal(alt["episode_starts"][:K, 0], np.array([r["start"] for r in rows], np.float32))
    assert carry["alt_last_done"][0] == float(latch)
    # ---- final ego observation (the obs SB3 would bootstrap from)
    if env_kind == "liar":
        st = carry["game_state"][0]
        obs_final = np.concatenate([st[:6], np.stack([st[12:24] & 7, st[12:24] >> 3], 1).reshape(-1)])
        n = int(st[24])
        want = g["final_obs"]
        assert np.array_equal(obs_final[:6 + 2 * n], want[:6 + 2 * n])
        assert np.all(want[6 + 2 * n::2] == 6) and np.all(want[7 + 2 * n::2] == 0)
    # episode statistics
    assert carry["ep_stats"][0] == g["ego_done"].sum()
    assert carry["ep_stats"][2] == T and carry["ep_stats"][3] == K


def test_rollout_split_equals_one_long_rollout():
    """Carry state across rollout boundaries: 2 x 32 ticks == 64 ticks for the ego AND for the
    partner: a partner row still waiting for the ego's next move when the first rollout ends is carried
    into the second one as its row 0, so no reward is lost at the boundary (the closed rows of the two
    short rollouts, put end to end, are the closed rows of the long one — rewards included)."""
    space = oracle.make_space(**oracle.LIAR_SPACE)
    P = oracle.param_count(space)
    pe = (np.random.RandomState(0).randn(P) * 0.3).astype(np.float32)
    pa = (np.random.RandomState(1).randn(P) * 0.3).astype(np.float32)
    N = 64
    e_all, a_all, c_all = orc.rollout("liar", space, pe, pa, N=N, T=64, seed=3)
    e1, a1, c1 = orc.rollout("liar", space, pe, pa, N=N, T=32, seed=3)
    cnt1 = a1["count"].copy()
    open1 = (c1["flags"] >> 2) & 1
    assert 0 < open1.sum() < N and np.all(c1["alt_last_done"][open1 == 1] == 0)
    a1 = {k: v.copy() for k, v in a1.items()}
    a2 = {k: v.copy() for k, v in a1.items()}  # the SAME buffer goes into the next rollout (the open row is in it)
    e2, a2, c2 = orc.rollout("liar", space, pe, pa, N=N, T=32, seed=3, tick0=32, first_rollout=False,
                             carry=c1, alt=a2)
    # the boundary case itself: a carried row that the ego's first move of the next rollout finishes with the
    # game's only reward (+-1) — the reward the per-rollout restart of the partner's buffer used to lose
    carried_terminal = [n for n in range(N) if open1[n] and a2["rewards"][0, n] != 0]
    assert carried_terminal and all(a2["episode_starts"][1, n] == 1 for n in carried_terminal if a2["count"][n] > 1)
    assert np.array_equal(c2["alt_boot_done"], c_all["alt_boot_done"])
    assert np.array_equal(c2["flags"], c_all["flags"])
    for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
        assert np.array_equal(np.concatenate([e1[k], e2[k]]), e_all[k]), k
    assert np.array_equal(c2["ego_last_value"], c_all["ego_last_value"])
    assert np.array_equal(cnt1 + a2["count"], a_all["count"])
    for n in range(N):
        k1, k2 = cnt1[n], a2["count"][n] + ((c2["flags"][n] >> 2) & 1)  # closed rows + the open one, if any
        for k in ("obs", "actions", "values", "logp", "episode_starts", "rewards"):
            assert np.array_equal(a_all[k][:k1, n], a1[k][:k1, n])
            assert np.array_equal(a_all[k][k1:k1 + k2, n], a2[k][:k2, n])


def test_rollout_statistics_are_sane():
    space = oracle.make_space(**oracle.LIAR_SPACE)
    P = oracle.param_count(space)
    pe = (np.random.RandomState(0).randn(P) * 0.1).astype(np.float32)
    ego, alt, carry = orc.rollout("liar", space, pe, pe, N=256, T=64, seed=1)
    assert set(np.unique(ego["rewards"])) <= {-1.0, 0.0, 1.0}
    done_next = np.concatenate([ego["episode_starts"][1:], carry["ego_last_done"][None]])
    assert np.all((ego["rewards"] != 0) <= (done_next == 1))  # rewards only on terminal ticks
    assert alt["count"].min() >= 1 and alt["count"].max() <= 128
    # zero-sum: partner rewards mirror the ego's
    assert abs(ego["rewards"].sum() + sum(alt["rewards"][:alt["count"][n], n].sum() for n in range(256))) <= 64
    rp = oracle.make_space(**oracle.RPS_SPACE)
    pr = (np.random.RandomState(2).randn(oracle.param_count(rp)) * 0.5).astype(np.float32)
    ego, alt, carry = orc.rollout("rps", rp, pr, pr, N=128, T=16, seed=2)
    assert np.all(ego["episode_starts"] == 1) and np.all(alt["count"] == 16)
    assert np.array_equal(ego["rewards"], -alt["rewards"][:16])
