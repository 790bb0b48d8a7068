"""GPU parity of the ADAP path (pantheonrl/algos/adap): AdapPolicy's forward (context inputs behind the features)
and ADAP.train (PPO's losses + the context KL loss as extra tiles) through the C ABI vs the CPU oracle, bit for
bit; and against the parameters the reference's own ADAP.train produced (tests/golden/adap.npz, tolerance)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import update as oupd
from pantheonrl_b200 import _lib, ops, update as dupd
from test_oracle_update import make_batch

pytestmark = pytest.mark.gpu
BOX = dict(box_dim=62, heads=[6])


def d(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def gspace(kw):
    return _lib.Space.box(kw["box_dim"], kw["heads"]) if "box_dim" in kw else _lib.Space.onehot(kw["nvec"], kw["heads"])


def rand_adap_params(osp, C, seed, scale=0.3):
    return (scale * np.random.RandomState(seed).randn(oracle.adap_param_count(osp, C))).astype(np.float32)


def batch(kw, M, seed):
    if "box_dim" in kw:
        rng = np.random.RandomState(seed)
        obs = np.zeros((M, 64), np.float32)
        obs[:, :62] = rng.randint(-4, 5, (M, 62))
        act = np.zeros((M, 4), np.uint8)
        act[:, 0] = rng.randint(0, 6, M)
        return obs, act, rng.randn(M).astype(np.float32), rng.randn(M).astype(np.float32)
    obs, act, _, adv, ret = make_batch(kw, M, seed=seed)
    return obs, act, adv, ret


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, BOX])
@pytest.mark.parametrize("C,B,per_sample", [(3, 1, False), (3, 777, True), (5, 300, False), (8, 129, True)])
def test_adap_forward_bit_exact(ctx, kw, C, B, per_sample):
    osp, sp = oracle.make_space(**kw), gspace(kw)
    params = rand_adap_params(osp, C, seed=B)
    obs, act, _, _ = batch(kw, B, B + 1)
    cx = np.random.RandomState(C).randn(B if per_sample else 1, C).astype(np.float32)
    for action_in in (None, act):
        want = oracle.adap_forward(osp, params, obs, cx if per_sample else cx[0], seed=3, tick=9, idx0=5,
                                   action_in=action_in)
        got = ops.policy_forward(sp, d(params), d(obs), seed=3, tick=9, idx0=5, context=d(cx),
                                 action_in=None if action_in is None else d(action_in))
        for k in ("action", "value", "logp", "entropy", "logits"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (k, action_in is None)


def run_both(kw, C, params, obs, act, old_logp, adv, ret, cx, perm, BS, grid, sidx=None, draws=None, coeff=0.0, **hp):
    osp, sp = oracle.make_space(**kw), gspace(kw)
    M = perm.shape[1]
    n = perm.shape[0] * (-(-M // BS))
    dp, dm, dv = d(params), d(np.zeros_like(params)), d(np.zeros_like(params))
    ws = dupd.UpdateWorkspace(sp, M, BS, context_size=C)
    cl = torch.full((n,), -1.0, device="cuda")
    extra = {} if sidx is None else dict(loss_kind=_lib.PTH_LOSS_ADAP, context_loss_coeff=coeff, ctx_states=d(sidx),
                                         ctx_draws=d(draws), ctx_loss=cl)
    gst = dupd.ppo_update(sp, dp, dm, dv, 0, d(obs), d(act), d(old_logp), d(adv), d(ret), d(perm), BS, ws,
                          grid_ctas=grid, context=d(cx), **extra, **hp)
    torch.cuda.synchronize()
    op, om, ov = params.copy(), np.zeros_like(params), np.zeros_like(params)
    oextra = {} if sidx is None else dict(loss_kind=2, ctx_loss_coeff=coeff, ctx_sidx=sidx, ctx_draws=draws)
    out = oupd.ppo_update(osp, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, ctx=cx, **oextra, **hp)
    gst = gst.cpu().numpy()
    assert np.array_equal(gst, out[0]), np.abs(gst - out[0]).max()
    assert np.array_equal(dm.cpu().numpy(), om) and np.array_equal(dv.cpu().numpy(), ov)
    assert np.array_equal(dp.cpu().numpy(), op), np.abs(dp.cpu().numpy() - op).max()
    if sidx is not None:
        assert np.array_equal(cl.cpu().numpy(), out[2])
    return op, gst, cl.cpu().numpy()


def random_draws(rs, n, K, S, C, M, BS):
    """th.randperm(B)[:S] per minibatch (-1 beyond a short last minibatch) and K unit-sphere contexts."""
    n_mb = -(-M // BS)
    sidx = np.full((n, S), -1, np.int32)
    for i in range(n):
        B = min(BS, M - (i % n_mb) * BS)
        sidx[i, :min(S, B)] = rs.permutation(B)[:S]
    dr = rs.rand(n, K, C).astype(np.float32) * 2 - 1
    dr = dr / np.sqrt((dr ** 2).sum(-1, keepdims=True)).astype(np.float32)
    return sidx, dr.astype(np.float32)


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, BOX])
@pytest.mark.parametrize("M,BS,E,grid", [(300, 300, 1, 2), (700, 256, 2, 3), (1000, 64, 2, 1)])
def test_adap_policy_update_bit_exact_vs_oracle(ctx, kw, M, BS, E, grid):
    """AdapPolicy under plain PPO.train (context inputs, no context loss): what a FIXED-context ADAP partner runs."""
    C = 3
    osp = oracle.make_space(**kw)
    params = rand_adap_params(osp, C, seed=M)
    obs, act, adv, ret = batch(kw, M, M)
    cx = np.random.RandomState(5).randn(M, C).astype(np.float32)
    ev = oracle.adap_forward(osp, params, obs, cx, action_in=act)
    old_logp = (ev["logp"] + 0.1 * np.random.RandomState(2).randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    run_both(kw, C, params, obs, act, old_logp, adv, ret, cx, perm, BS, grid, ent_coef=0.01)


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, BOX])
@pytest.mark.parametrize("M,BS,E,grid,K,S,C", [(280, 64, 2, 1, 5, 32, 3), (700, 256, 2, 3, 5, 32, 3),
                                             (600, 300, 1, 96, 3, 20, 2), (500, 250, 2, 4, 16, 40, 8),
                                             (1200, 600, 1, 2, 2, 100, 3)])
def test_adap_train_bit_exact_vs_oracle(ctx, kw, M, BS, E, grid, K, S, C):
    """ADAP.train: context tiles behind the minibatch's own tiles (several per minibatch when S > 128 / K,
    a short last one, more CTAs than tiles, CTAs whose only tile is a context tile)."""
    osp = oracle.make_space(**kw)
    params = rand_adap_params(osp, C, seed=M + K)
    obs, act, adv, ret = batch(kw, M, M)
    rs = np.random.RandomState(K)
    cx = rs.randn(M, C).astype(np.float32)
    ev = oracle.adap_forward(osp, params, obs, cx, action_in=act)
    old_logp = (ev["logp"] + 0.1 * rs.randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    sidx, draws = random_draws(rs, E * (-(-M // BS)), K, S, C, M, BS)
    _, st, cl = run_both(kw, C, params, obs, act, old_logp, adv, ret, cx, perm, BS, grid, sidx, draws, coeff=0.7,
                         ent_coef=0.01)
    assert np.all((cl > 0) & (cl <= 1.0 + 1e-6)) and np.all(np.isfinite(st))


@pytest.mark.parametrize("name,kw", [("rps", oracle.RPS_SPACE), ("liar", oracle.LIAR_SPACE), ("liar_k3", oracle.LIAR_SPACE)])
def test_adap_train_reproduces_the_reference_run(ctx, name, kw):
    """The inputs and random draws of the reference's own ADAP.train run (tests/golden/make_golden_adap.py) through
    the CUDA kernel: bit-exact vs the oracle, and within fp32 tolerance of the parameters the reference ended with."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "adap.npz"))
    pre = name + "_"
    M, BS, E, K, S = (int(x) for x in g[pre + "hp"])
    log = dict(zip(g[pre + "log_keys"], g[pre + "log_vals"]))
    n_mb = -(-M // BS)
    p, st, cl = run_both(kw, 3, g[pre + "p0"], g[pre + "obs"], g[pre + "act"], g[pre + "old_logp"], g[pre + "adv"],
                         g[pre + "ret"], g[pre + "ctx"], g[pre + "perms"], BS, 3, g[pre + "sidx"], g[pre + "draws"],
                         coeff=float(g[pre + "coeff"][0]), ent_coef=0.01)
    assert np.abs(p - g[pre + "params"]).max() <= 5e-6
    assert cl[-n_mb:].mean() == pytest.approx(log["train/context_kl_loss"], abs=1e-5)
    assert st[-1, 5] == pytest.approx(log["train/loss"], abs=2e-5)


@pytest.mark.parametrize("sampler,C", [("l2", 3), ("l2", 8), ("unit_square", 5), ("positive_square", 2), ("categorical", 4),
                                       ("natural_numbers", 1)])
def test_adap_draws_bit_exact_vs_oracle(ctx, sampler, C):
    """pth_adap_draw: the Philox stand-in for SAMPLERS / th.randperm (adap/util.py:42-106)."""
    n, K, S, n_mb, M, BS = 11, 5, 32, 4, 210, 64  # the last minibatch of an epoch holds 18 < S samples
    want_s, want_d = oupd.adap_draw(n, K, C, sampler, 10, 0x10300, index0=7, S=S, n_mb=n_mb, M=M, batch_size=BS)
    got_s, got_d = dupd.adap_draw(n, K, C, sampler, 10, 0x10300, index0=7, S=S, n_mb=n_mb, M=M, batch_size=BS)
    assert np.array_equal(got_s.cpu().numpy(), want_s) and np.array_equal(got_d.cpu().numpy(), want_d)
    for i in range(n):
        B = min(BS, M - (i % n_mb) * BS)
        v = want_s[i][want_s[i] >= 0]
        assert len(v) == min(S, B) and len(set(v.tolist())) == len(v) and v.max() < B
    if sampler == "l2":
        assert np.allclose(np.linalg.norm(want_d, axis=-1), 1.0, atol=1e-6)
    _, one = dupd.adap_draw(1, 1, C, sampler, 10, 0x10300, index0=8)  # a single per-episode context
    assert one.shape == (1, 1, C)


# ------------------------------------------------------------------ AdapPolicyMult (ADAP_MULT)
def rand_mult_params(osp, C, seed, scale=0.3):
    return (scale * np.random.RandomState(seed).randn(oracle.adap_mult_param_count(osp, C))).astype(np.float32)


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, BOX])
@pytest.mark.parametrize("C,B,per_sample", [(3, 1, False), (3, 777, True), (8, 129, True)])
def test_adap_mult_forward_bit_exact(ctx, kw, C, B, per_sample):
    osp, sp = oracle.make_space(**kw), gspace(kw)
    params = rand_mult_params(osp, C, seed=B)
    assert params.size == _lib.load().pth_adap_mult_param_count(sp, C)
    obs, act, _, _ = batch(kw, B, B + 1)
    cx = np.random.RandomState(C).randn(B if per_sample else 1, C).astype(np.float32)
    for action_in in (None, act):
        want = oracle.adap_mult_forward(osp, params, obs, cx if per_sample else cx[0], seed=3, tick=9, idx0=5,
                                        action_in=action_in)
        got = ops.policy_forward(sp, d(params), d(obs), seed=3, tick=9, idx0=5, context=d(cx), adap_mult=True,
                                 action_in=None if action_in is None else d(action_in))
        for k in ("action", "value", "logp", "entropy", "logits"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (k, action_in is None)


def run_both_mult(kw, C, params, obs, act, old_logp, adv, ret, cx, perm, BS, grid, sidx=None, draws=None, coeff=0.0, **hp):
    osp, sp = oracle.make_space(**kw), gspace(kw)
    M = perm.shape[1]
    n = perm.shape[0] * (-(-M // BS))
    dp, dm, dv = d(params), d(np.zeros_like(params)), d(np.zeros_like(params))
    ws = dupd.UpdateWorkspace(sp, M, BS, context_size=C, adap_mult=True)
    cl = torch.full((n,), -1.0, device="cuda")
    extra = {} if sidx is None else dict(loss_kind=_lib.PTH_LOSS_ADAP, context_loss_coeff=coeff, ctx_states=d(sidx),
                                         ctx_draws=d(draws), ctx_loss=cl)
    gst = dupd.ppo_update(sp, dp, dm, dv, 0, d(obs), d(act), d(old_logp), d(adv), d(ret), d(perm), BS, ws,
                          grid_ctas=grid, context=d(cx), adap_mult=True, **extra, **hp)
    torch.cuda.synchronize()
    op, om, ov = params.copy(), np.zeros_like(params), np.zeros_like(params)
    oextra = {} if sidx is None else dict(loss_kind=2, ctx_loss_coeff=coeff, ctx_sidx=sidx, ctx_draws=draws)
    out = oupd.ppo_update(osp, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, ctx=cx, adap_mult=True,
                          **oextra, **hp)
    gst = gst.cpu().numpy()
    assert np.array_equal(gst, out[0]), np.abs(gst - out[0]).max()
    assert np.array_equal(dm.cpu().numpy(), om) and np.array_equal(dv.cpu().numpy(), ov)
    assert np.array_equal(dp.cpu().numpy(), op), np.abs(dp.cpu().numpy() - op).max()
    if sidx is not None:
        assert np.array_equal(cl.cpu().numpy(), out[2])
    return op, gst, cl.cpu().numpy()


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, BOX])
@pytest.mark.parametrize("M,BS,E,grid,K,S,C", [(280, 64, 2, 1, 5, 32, 3), (700, 256, 2, 3, 0, 0, 3), (600, 300, 1, 96, 3, 20, 2),
                                             (500, 250, 1, 4, 16, 40, 8)])
def test_adap_mult_train_bit_exact_vs_oracle(ctx, kw, M, BS, E, grid, K, S, C):
    """AdapPolicyMult under PPO.train (K = 0) and ADAP.train (context tiles), several tiles per CTA, 96 CTAs."""
    osp = oracle.make_space(**kw)
    params = rand_mult_params(osp, C, seed=M + K)
    obs, act, adv, ret = batch(kw, M, M)
    rs = np.random.RandomState(K + 1)
    cx = rs.randn(M, C).astype(np.float32)
    ev = oracle.adap_mult_forward(osp, params, obs, cx, action_in=act)
    old_logp = (ev["logp"] + 0.1 * rs.randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    if K:
        sidx, draws = random_draws(rs, E * (-(-M // BS)), K, S, C, M, BS)
        run_both_mult(kw, C, params, obs, act, old_logp, adv, ret, cx, perm, BS, grid, sidx, draws, coeff=0.7, ent_coef=0.01)
    else:
        run_both_mult(kw, C, params, obs, act, old_logp, adv, ret, cx, perm, BS, grid, ent_coef=0.01)


@pytest.mark.parametrize("name,kw", [("mult_rps", oracle.RPS_SPACE), ("mult_liar", oracle.LIAR_SPACE)])
def test_adap_mult_train_reproduces_the_reference_run(ctx, name, kw):
    """The reference's own MultModel + ADAP.train run (tests/golden/make_golden_adap.py) through the CUDA kernel."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "adap.npz"))
    pre = name + "_"
    M, BS, E, K, S = (int(x) for x in g[pre + "hp"])
    log = dict(zip(g[pre + "log_keys"], g[pre + "log_vals"]))
    n_mb = -(-M // BS)
    p, st, cl = run_both_mult(kw, 3, g[pre + "p0"], g[pre + "obs"], g[pre + "act"], g[pre + "old_logp"], g[pre + "adv"],
                              g[pre + "ret"], g[pre + "ctx"], g[pre + "perms"], BS, 3, g[pre + "sidx"], g[pre + "draws"],
                              coeff=float(g[pre + "coeff"][0]), ent_coef=0.01)
    assert np.abs(p - g[pre + "params"]).max() <= 2e-6
    assert cl[-n_mb:].mean() == pytest.approx(log["train/context_kl_loss"], abs=1e-5)
    assert st[-1, 5] == pytest.approx(log["train/loss"], abs=2e-5)
