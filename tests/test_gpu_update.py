"""GPU parity: the fused PPO update kernel vs the CPU oracle (bit-exact: same
reduction tree, same Adam) and vs an independent torch-autograd restatement of
SB3 PPO.train (fp32 tolerance), plus the shuffle / compaction helpers."""
import numpy as np
import pytest
import torch

import oracle
from oracle import sb3_torch
from oracle import update as oupd
from pantheonrl_b200 import _lib, update as dupd
from test_oracle_update import make_batch

pytestmark = pytest.mark.gpu


def gspace(kw):
    return _lib.Space.onehot(kw["nvec"], kw["heads"])


def run_gpu(kw, params, m, v, step, obs, act, old_logp, adv, ret, perm, BS, grid, index=None, **hp):
    sp = gspace(kw)
    M = perm.shape[1]
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
    dp, dm, dv = d(params), d(m), d(v)
    ws = dupd.UpdateWorkspace(sp, M, BS)
    stats = dupd.ppo_update(sp, dp, dm, dv, step, d(obs), d(act), d(old_logp), d(adv), d(ret), d(perm),
                            BS, ws, index=None if index is None else d(index), grid_ctas=grid, **hp)
    torch.cuda.synchronize()
    return dp.cpu().numpy(), dm.cpu().numpy(), dv.cpu().numpy(), stats.cpu().numpy()


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, oracle.LIAR3_SPACE])
@pytest.mark.parametrize("M,BS,E,grid", [(300, 300, 1, 2), (700, 256, 3, 3), (2048, 512, 2, 4), (1000, 64, 2, 1)])
def test_update_bit_exact_vs_oracle(ctx, kw, M, BS, E, grid):
    space = oracle.make_space(**kw)
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=M)
    params = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=M)
    ev = oracle.policy_forward(space, params, obs, action_in=act)
    old_logp = (ev["logp"] + 0.1 * np.random.RandomState(2).randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    m, v = np.zeros_like(params), np.zeros_like(params)
    hp = dict(ent_coef=0.01)
    gp, gm, gv, gst = run_gpu(kw, params, m, v, 0, obs, act, old_logp, adv, ret, perm, BS, grid, **hp)
    op, om, ov = params.copy(), m.copy(), v.copy()
    ost, _ = oupd.ppo_update(space, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, **hp)
    assert np.array_equal(gst, ost), np.abs(gst - ost).max()
    assert np.array_equal(gm, om) and np.array_equal(gv, ov)
    assert np.array_equal(gp, op), np.abs(gp - op).max()


@pytest.mark.parametrize("grid", [72, 140])
def test_update_bit_exact_with_many_ctas(ctx, grid):
    """One tile per CTA on 72 / 140 CTAs: the ordered reduction then splits every parameter's chain
    over 2 / 4 thread groups (<= 64 partials is the one-group path the small-grid tests cover)."""
    kw = oracle.RPS_SPACE
    M = BS = 128 * grid
    space = oracle.make_space(**kw)
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=grid)
    params = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=grid)
    ev = oracle.policy_forward(space, params, obs, action_in=act)
    old_logp = (ev["logp"] + 0.1 * np.random.RandomState(2).randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, 2, seed=10, stream=4)
    m, v = np.zeros_like(params), np.zeros_like(params)
    gp, gm, gv, gst = run_gpu(kw, params, m, v, 0, obs, act, old_logp, adv, ret, perm, BS, grid, ent_coef=0.01)
    op, om, ov = params.copy(), m.copy(), v.copy()
    ost, _ = oupd.ppo_update(space, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, ent_coef=0.01)
    assert np.array_equal(gst, ost) and np.array_equal(gm, om) and np.array_equal(gv, ov) and np.array_equal(gp, op)


def test_subnormal_partial_sums_follow_red_add_semantics(ctx):
    """Later tiles of a CTA are added with RED.ADD.F32, which flushes subnormal inputs and
    results to zero; the oracle restates that.  Advantages of ~1e-33 with vf_coef = ent_coef = 0
    put the policy-side gradient partials around FLT_MIN: some normal, some subnormal (and there
    are several tiles per CTA)."""
    kw = oracle.LIAR_SPACE
    M, BS, E, grid = 1024, 1024, 1, 2  # 8 tiles on 2 CTAs
    space = oracle.make_space(**kw)
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=11)
    params = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=11)
    adv = (adv * np.float32(1e-33)).astype(np.float32)
    ev = oracle.policy_forward(space, params, obs, action_in=act)
    old_logp = ev["logp"].astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    m, v = np.zeros_like(params), np.zeros_like(params)
    hp = dict(ent_coef=0.0, vf_coef=0.0, normalize_advantage=False)
    gp, gm, gv, gst = run_gpu(kw, params, m, v, 0, obs, act, old_logp, adv, ret, perm, BS, grid, **hp)
    op, om, ov = params.copy(), m.copy(), v.copy()
    ost, _ = oupd.ppo_update(space, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, **hp)
    assert np.array_equal(gm, om) and np.array_equal(gv, ov)
    assert np.array_equal(gp, op) and np.array_equal(gst, ost)
    tiny = np.abs(gm[gm != 0])
    assert tiny.size and tiny.min() < 1e-36  # the regime was reached: first moments are that small


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE])
@pytest.mark.parametrize("M,BS,E,grid,l2", [(150, 32, 2, 3, 0.0), (700, 300, 2, 2, 0.01)])
def test_behaviour_cloning_bit_exact_vs_oracle(ctx, kw, M, BS, E, grid, l2):
    """loss_kind = PTH_LOSS_BC (pantheonrl/algos/bc.py:270-315): same kernel, supervised loss."""
    space = oracle.make_space(**kw)
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=M)
    params = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=M + 1)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    m, v = np.zeros_like(params), np.zeros_like(params)
    hp = dict(loss_kind=1, l2_weight=l2, ent_coef=1e-3, vf_coef=0.0, max_grad_norm=float("inf"),
              learning_rate=1e-3, eps=1e-8, normalize_advantage=False)
    gp, gm, gv, gst = run_gpu(kw, params, m, v, 0, obs, act, old_logp, adv, ret, perm, BS, grid, **hp)
    op, om, ov = params.copy(), m.copy(), v.copy()
    ost, _ = oupd.ppo_update(space, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, **hp)
    assert np.array_equal(gst, ost), np.abs(gst - ost).max()
    assert np.array_equal(gm, om) and np.array_equal(gv, ov) and np.array_equal(gp, op)
    assert np.all(gst[:, 4] == 0) and np.all(np.isfinite(gst))  # (random labels: nothing to learn here)


def test_update_matches_torch_autograd(ctx):
    kw = oracle.LIAR_SPACE
    M, BS, E = 1500, 512, 2
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=5)
    space = oracle.make_space(**kw)
    params = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=1)
    ev = oracle.policy_forward(space, params, obs, action_in=act)
    old_logp = (ev["logp"] + 0.05 * np.random.RandomState(3).randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=1, stream=5)
    m, v = np.zeros_like(params), np.zeros_like(params)
    gp, gm, gv, gst = run_gpu(kw, params, m, v, 0, obs, act, old_logp, adv, ret, perm, BS, 0)
    ref = sb3_torch.ppo_train(pol, obs[:, :30], act[:, :2], old_logp, adv, ret, perm, BS)
    assert np.allclose(gp, pol.to_flat(), atol=5e-6, rtol=0)
    for i, k in enumerate(("pg_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction", "loss", "grad_norm")):
        want = np.array([r[k] for r in ref])
        assert np.allclose(gst[:, i], want, atol=5e-5, rtol=1e-4), k


def test_update_with_index_and_adam_state(ctx):
    kw = oracle.RPS_SPACE
    space = oracle.make_space(**kw)
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=2)
    params = pol.to_flat().copy()
    Mall, M = 900, 300
    obs, act, old_logp, adv, ret = make_batch(kw, Mall, seed=4)
    index = (np.arange(M) * 3 + 2).astype(np.int32)
    perm = oupd.perm_feistel(M, 2, seed=3, stream=4)
    rng = np.random.RandomState(0)
    m = (rng.randn(params.size) * 1e-3).astype(np.float32)
    v = (rng.rand(params.size) * 1e-5).astype(np.float32)
    gp, gm, gv, gst = run_gpu(kw, params, m, v, 37, obs, act, old_logp, adv, ret, perm, 128, 2, index=index,
                              normalize_advantage=False)
    op, om, ov = params.copy(), m.copy(), v.copy()
    ost, _ = oupd.ppo_update(space, op, om, ov, 37, obs, act, old_logp, adv, ret, perm, 128, 2, index=index,
                             normalize_advantage=False)
    assert np.array_equal(gp, op) and np.array_equal(gm, om) and np.array_equal(gv, ov)
    assert np.array_equal(gst, ost)


def test_auto_grid_and_full_size_run(ctx):
    # BASELINE config 2 shape: 128 x 4096 samples, 32 minibatches per epoch
    kw = oracle.LIAR_SPACE
    sp = gspace(kw)
    M, BS = 128 * 4096, 128 * 4096 // 32
    G = dupd.update_grid(sp, M, BS)
    assert 1 <= G <= ctx.sm_count and G == min(ctx.sm_count, BS // 128)
    # one-tile minibatches (the reference's n_envs = 1, batch_size = 64) still spread reduction + Adam
    assert dupd.update_grid(sp, 2048, 64) == min(ctx.sm_count, 96)
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=1)
    params = torch.from_numpy(pol.to_flat().copy()).cuda()
    p0 = params.clone()
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.zeros(M, 32, dtype=torch.uint8, device="cuda")
    for s, n in enumerate(kw["nvec"]):
        obs[:, s] = torch.randint(0, n, (M,), generator=g, device="cuda", dtype=torch.uint8)
    act = torch.zeros(M, 4, dtype=torch.uint8, device="cuda")
    act[:, 0] = torch.randint(0, 7, (M,), generator=g, device="cuda", dtype=torch.uint8)
    act[:, 1] = torch.randint(0, 12, (M,), generator=g, device="cuda", dtype=torch.uint8)
    from pantheonrl_b200 import ops
    old_logp = ops.policy_forward(sp, params, obs, action_in=act, want=("logp",))["logp"]
    adv = torch.randn(M, generator=g, device="cuda")
    ret = torch.randn(M, generator=g, device="cuda")
    perm = dupd.perm_feistel(M, 2, seed=10, stream_id=4)
    assert torch.equal(torch.sort(perm[0]).values, torch.arange(M, device="cuda", dtype=torch.int32))
    ws = dupd.UpdateWorkspace(sp, M, BS)
    m, v = torch.zeros_like(params), torch.zeros_like(params)
    st = dupd.ppo_update(sp, params, m, v, 0, obs, act, old_logp, adv, ret, perm, BS, ws)
    torch.cuda.synchronize()
    st = st.cpu().numpy()
    assert np.all(np.isfinite(st)) and np.all(st[:, 7] == BS)
    assert st[0, 3] == pytest.approx(0.0, abs=1e-6)      # first minibatch: ratio == 1 exactly
    assert st[0, 4] == 0.0
    assert float((params - p0).abs().max()) > 1e-5 and bool(torch.isfinite(params).all())
    # run-to-run determinism of the whole train()
    params2, m2, v2 = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    st2 = dupd.ppo_update(sp, params2, m2, v2, 0, obs, act, old_logp, adv, ret, perm, BS, ws)
    torch.cuda.synchronize()
    assert torch.equal(params, params2) and np.array_equal(st, st2.cpu().numpy())


@pytest.mark.parametrize("M", [1, 5, 1000, 70000])
def test_feistel_matches_oracle(ctx, M):
    want = oupd.perm_feistel(M, 3, seed=10, stream=4, epoch0=7)
    got = dupd.perm_feistel(M, 3, seed=10, stream_id=4, epoch0=7)
    assert np.array_equal(got.cpu().numpy(), want)


def test_index_build_matches_oracle(ctx):
    T, N = 37, 5000
    count = torch.randint(0, T + 1, (N,), generator=torch.Generator().manual_seed(0), dtype=torch.int32)
    want = oupd.index_build(count.numpy(), T, N)
    idx, total = dupd.index_build(count.cuda(), T, N)
    assert int(total.item()) == want.size
    assert np.array_equal(idx[:want.size].cpu().numpy(), want)
    idx, total = dupd.index_build(None, 3, 4)
    assert idx.cpu().numpy().tolist() == oupd.index_build(None, 3, 4).tolist()
