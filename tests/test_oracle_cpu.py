"""CPU tests: pin the oracle against golden vectors generated from the
reference's own classes (tests/golden/make_golden.py), known-answer vectors,
closed forms and an independent numpy restatement."""
import math
import os

import numpy as np
import pytest

import oracle
from oracle import sb3_numpy


# ----------------------------------------------------------------- RNG
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        got = oracle.philox_raw(ctr, key)
        assert [int(x) for x in got] == want


def test_philox_stream_keying():
    a = oracle.philox(10, 1, 5, 7, 0)
    b = oracle.philox(10, 2, 5, 7, 0)
    c = oracle.philox(10, 1, 5, 7, 1)
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    # index_hi lands in counter word 3, the seed's high half in key word 1
    d = oracle.philox_raw([5, 7, 0, 1], [(10 ^ (1 * 0x9E3779B9)) & 0xffffffff, 3])
    assert np.array_equal(d, oracle.philox(10 | (3 << 32), 1, 5 | (1 << 32), 7, 0))


# ----------------------------------------------------------------- math
def _ulp_err(got, want64):
    got = got.astype(np.float64)
    ulp = np.spacing(np.abs(want64).astype(np.float32)).astype(np.float64)
    return np.abs(got - want64) / ulp


def test_expf_accuracy():
    x = np.concatenate([np.linspace(-87, 88, 200001), np.linspace(-1, 1, 100001)]).astype(np.float32)
    err = _ulp_err(oracle.math_vec("exp", x), np.exp(x.astype(np.float64)))
    assert err.max() <= 2.0
    assert oracle.math_vec("exp", [-100.0])[0] == 0.0
    assert oracle.math_vec("exp", [0.0])[0] == 1.0


def test_logf_accuracy():
    x = np.concatenate([np.linspace(1, 40, 200001), np.geomspace(1e-30, 1e30, 100001)]).astype(np.float32)
    err = _ulp_err(oracle.math_vec("log", x), np.log(x.astype(np.float64)))
    assert err.max() <= 2.0
    assert oracle.math_vec("log", [1.0])[0] == 0.0


def test_tanhf_accuracy():
    x = np.concatenate([np.linspace(-12, 12, 400001), np.linspace(-0.7, 0.7, 100001)]).astype(np.float32)
    err = _ulp_err(oracle.math_vec("tanh", x), np.tanh(x.astype(np.float64)))
    assert err.max() <= 3.0
    y = oracle.math_vec("tanh", x)
    assert np.all(np.abs(y) <= 1.0)
    assert np.array_equal(oracle.math_vec("tanh", -x), -y)  # odd


# ----------------------------------------------------------------- GAE
def _gae_inputs(T, N, p_start, seed=0):
    rng = np.random.RandomState(seed)
    rew = rng.randint(-1, 2, size=(T, N)).astype(np.float32)
    val = rng.randn(T, N).astype(np.float32)
    start = (rng.rand(T, N) < p_start).astype(np.float32)
    lv = rng.randn(N).astype(np.float32)
    dn = (rng.rand(N) < p_start).astype(np.float32)
    return rew, val, start, lv, dn


@pytest.mark.parametrize("T,N,p", [(2048, 1, 1.0), (128, 37, 0.25), (400, 64, 1 / 400), (1, 5, 0.5)])
def test_gae_c_oracle_matches_numpy_restatement(T, N, p):
    rew, val, start, lv, dn = _gae_inputs(T, N, p)
    a0, r0 = sb3_numpy.compute_returns_and_advantage(rew, val, start, lv, dn)
    a1, r1 = oracle.gae(rew, val, start, lv, dn)
    # same fp32 ops in the same order: bit-identical
    assert np.array_equal(a0, a1) and np.array_equal(r0, r1)


def test_gae_closed_forms():
    T, N = 50, 3
    g, lam = 0.99, 0.95
    # every step terminal (RPS): A_t = r_t - V_t exactly
    rew, val, _, lv, _ = _gae_inputs(T, N, 1.0, seed=3)
    a, r = oracle.gae(rew, val, np.ones((T, N)), lv, np.ones(N), g, lam)
    assert np.array_equal(a, rew - val) and np.array_equal(r, (rew - val) + val)
    # no terminals, V = 0, r = 1: A_t = sum_k (g*lam)^k
    a, _ = oracle.gae(np.ones((T, N)), np.zeros((T, N)), np.zeros((T, N)), np.zeros(N), np.zeros(N), g, lam)
    c = g * lam
    want = np.array([(1 - c ** (T - t)) / (1 - c) for t in range(T)])
    assert np.allclose(a[:, 0], want, rtol=0, atol=1e-5)


def test_gae_ragged_reduces_to_dense_with_quirk():
    T, N = 40, 9
    rew, val, start, _, dn = _gae_inputs(T, N, 0.3, seed=5)
    count = np.random.RandomState(1).randint(0, T + 1, size=N).astype(np.int32)
    count[0] = T
    count[1] = 0
    a, r = oracle.gae_ragged(rew, val, start, count, dn)
    for n in range(N):
        c = int(count[n])
        if c == 0:
            assert not a[:, n].any()
            continue
        # agents.py:127-129: bootstrap value = value of the last stored step
        ad, rd = sb3_numpy.compute_returns_and_advantage(
            rew[:c, n:n + 1], val[:c, n:n + 1], start[:c, n:n + 1], val[c - 1, n:n + 1], dn[n:n + 1])
        assert np.array_equal(a[:c, n], ad[:, 0]) and np.array_equal(r[:c, n], rd[:, 0])
        assert not a[c:, n].any()


# ----------------------------------------------------------------- games vs the reference's own classes
def test_rps_payoff_matches_reference(golden_dir):
    tab = np.load(os.path.join(golden_dir, "rps_payoff.npz"))["table"]
    re, ra = oracle.rps_step(tab[:, 0], tab[:, 1])
    assert np.array_equal(re, tab[:, 2].astype(np.float32))
    assert np.array_equal(ra, tab[:, 3].astype(np.float32))
    assert np.all(tab[:, 4] == 1)


def _state_from_hands(hands):
    s = np.zeros((hands.shape[0], 32), np.uint8)
    s[:, :12] = hands
    return s


def replay_liar_env(g, step_fn):
    """Replay the golden env-level trace through step_fn(state, is_ego, action)."""
    eps = g["ep_start"]
    n_ep = len(eps) - 1
    state = _state_from_hands(g["hands"])
    alive = np.ones(n_ep, bool)
    k = 0
    while alive.any():
        idx = np.where(alive)[0]
        pos = eps[idx] + k
        st = np.ascontiguousarray(state[idx])
        obs, re, ra, done = step_fn(st, g["is_ego"][pos], g["raw_action"][pos])
        state[idx] = st
        assert np.array_equal(obs[:, :30], g["obs"][pos]), f"obs mismatch at move {k}"
        assert np.array_equal(re, g["r_ego"][pos]) and np.array_equal(ra, g["r_alt"][pos])
        assert np.array_equal(done, g["done"][pos])
        k += 1
        alive = (eps[:-1] + k) < eps[1:]
        assert np.array_equal(done.astype(bool), ~alive[idx])


def test_liar_env_matches_reference(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "liar_env.npz")))
    replay_liar_env(g, oracle.liar_step)


def test_liar_forced_bluff_after_twelve_bids():
    # hand-derived: counts 0..11 strictly increasing fill the history; the 13th
    # move is coerced to a bluff call whatever the raw action (liar.py:60-62).
    hands = np.array([[1, 1, 1, 1, 1, 1, 6, 0, 0, 0, 0, 0]], np.uint8)
    st = _state_from_hands(hands)
    ego = 1
    for c in range(12):
        obs, re, ra, d = oracle.liar_step(st, [ego], [[0, c]])
        assert d[0] == 0 and re[0] == 0
        assert st[0, 24] == c + 1 and obs[0, 6] == 0 and obs[0, 7] == c
        ego ^= 1
    obs, re, ra, d = oracle.liar_step(st, [ego], [[3, 11]])
    # last bid: face 0, count index 11 -> claims 12 dice; truth: 1 + 6 = 7 -> a bluff.
    # caller is the ego (ego == 1 after 12 alternations): ego wins.
    assert d[0] == 1 and ego == 1 and re[0] == 1.0 and ra[0] == -1.0


def test_liar_reset_distribution_and_determinism():
    s1, ef1, o1 = oracle.liar_reset(20000, seed=10, tick=3)
    s2, ef2, o2 = oracle.liar_reset(20000, seed=10, tick=3)
    assert np.array_equal(s1, s2) and np.array_equal(ef1, ef2)
    assert np.all(s1[:, :6].sum(1) == 6) and np.all(s1[:, 6:12].sum(1) == 6)
    assert abs(ef1.mean() - 0.5) < 0.02
    face_freq = s1[:, :12].reshape(-1, 2, 6).sum((0, 1)) / (20000 * 12)
    assert np.all(np.abs(face_freq - 1 / 6) < 0.01)
    # obs is the mover's view of an empty table
    assert np.all(o1[:, 6:30:2] == 6) and np.all(o1[:, 7:30:2] == 0)
    mover_hand = np.where(ef1[:, None] == 1, s1[:, :6], s1[:, 6:12])
    assert np.array_equal(o1[:, :6], mover_hand)
    # shards: env0 offsets index the same global stream
    s3, _, _ = oracle.liar_reset(100, seed=10, tick=3, env0=500)
    assert np.array_equal(s3, s1[500:600])


# ----------------------------------------------------------------- policy forward (C oracle vs float64 math)
def _ref_forward64(space_kw, params, obs):
    nvec, heads = space_kw["nvec"], space_kw["heads"]
    F, L, H = sum(nvec), sum(heads), 64
    p = params.astype(np.float64)
    o = 0

    def take(n, shape):
        nonlocal o
        w = p[o:o + n].reshape(shape)
        o += n
        return w
    w_pi0, b_pi0 = take(H * F, (F, H)).T, take(H, (H,))  # stored input-major
    w_pi1, b_pi1 = take(H * H, (H, H)), take(H, (H,))
    w_vf0, b_vf0 = take(H * F, (F, H)).T, take(H, (H,))
    w_vf1, b_vf1 = take(H * H, (H, H)), take(H, (H,))
    w_act, b_act = take(L * H, (L, H)), take(L, (L,))
    w_val, b_val = take(H, (1, H)), take(1, (1,))
    B = obs.shape[0]
    x = np.zeros((B, F))
    off = 0
    for s, n in enumerate(nvec):
        x[np.arange(B), off + obs[:, s].astype(np.int64)] = 1.0
        off += n
    hp = np.tanh(np.tanh(x @ w_pi0.T + b_pi0) @ w_pi1.T + b_pi1)
    hv = np.tanh(np.tanh(x @ w_vf0.T + b_vf0) @ w_vf1.T + b_vf1)
    return hp @ w_act.T + b_act, (hv @ w_val.T + b_val)[:, 0]


def rand_params(space, seed=0, scale=0.3):
    P = oracle.param_count(space)
    return (np.random.RandomState(seed).randn(P) * scale).astype(np.float32)


def rand_liar_obs(B, seed=0):
    rng = np.random.RandomState(seed)
    obs = np.zeros((B, 32), np.uint8)
    for s, n in enumerate(oracle.LIAR_NVEC):
        obs[:, s] = rng.randint(n, size=B)
    return obs


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE])
def test_policy_forward_matches_float64(kw):
    space = oracle.make_space(**kw)
    assert oracle.param_count(space) == {1: 8836, 30: 44308}[len(kw["nvec"])]
    params = rand_params(space)
    B = 257
    obs = rand_liar_obs(B) if len(kw["nvec"]) == 30 else np.zeros((B, 32), np.uint8)
    out = oracle.policy_forward(space, params, obs, seed=10, tick=4)
    logits64, value64 = _ref_forward64(kw, params, obs)
    assert np.allclose(out["logits"], logits64, atol=2e-5, rtol=0)
    assert np.allclose(out["value"], value64, atol=2e-5, rtol=0)
    # log-prob / entropy against float64 log-softmax
    off = 0
    lp = np.zeros(B)
    en = np.zeros(B)
    for h, n in enumerate(kw["heads"]):
        z = logits64[:, off:off + n]
        ls = z - np.log(np.exp(z - z.max(1, keepdims=True)).sum(1, keepdims=True)) - z.max(1, keepdims=True)
        lp += ls[np.arange(B), out["action"][:, h]]
        en += -(np.exp(ls) * ls).sum(1)
        off += n
    assert np.allclose(out["logp"], lp, atol=2e-5, rtol=0)
    assert np.allclose(out["entropy"], en, atol=2e-5, rtol=0)
    # evaluate_actions path reproduces the sampled log-prob bit for bit
    ev = oracle.policy_forward(space, params, obs, action_in=out["action"])
    assert np.array_equal(ev["logp"], out["logp"]) and np.array_equal(ev["value"], out["value"])


def test_sampling_follows_the_distribution():
    space = oracle.make_space(**oracle.RPS_SPACE)
    params = rand_params(space, seed=4, scale=1.0)
    B = 60000
    obs = np.zeros((B, 32), np.uint8)
    out = oracle.policy_forward(space, params, obs, seed=7, tick=1)
    p = np.exp(out["logits"][0].astype(np.float64))
    p /= p.sum()
    freq = np.bincount(out["action"][:, 0], minlength=3) / B
    assert np.all(np.abs(freq - p) < 4 * np.sqrt(p * (1 - p) / B) + 1e-3)
