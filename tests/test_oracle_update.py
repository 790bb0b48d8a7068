"""CPU tests: the C oracle's hand-written PPO backward / clip / Adam against an
independent torch-autograd restatement of SB3 PPO.train (oracle/sb3_torch.py),
plus the shuffle / compaction helpers."""
import numpy as np
import pytest
import torch

import oracle
from oracle import sb3_torch
from oracle import update as oupd
from test_oracle_cpu import rand_liar_obs


def make_batch(kw, M, seed=0):
    rng = np.random.RandomState(seed)
    if len(kw["nvec"]) == 90:  # three stacked Liar's Dice frames in a 96-byte row
        obs = np.zeros((M, 96), np.uint8)
        for f in range(3):
            obs[:, 30 * f:30 * f + 30] = rand_liar_obs(M, seed + 17 * f)[:, :30]
    else:
        obs = rand_liar_obs(M, seed) if len(kw["nvec"]) == 30 else np.zeros((M, 32), np.uint8)
    act = np.zeros((M, 4), np.uint8)
    for h, n in enumerate(kw["heads"]):
        act[:, h] = rng.randint(n, size=M)
    old_logp = (-np.log(np.prod(kw["heads"])) + 0.1 * rng.randn(M)).astype(np.float32)
    adv = rng.randn(M).astype(np.float32)
    ret = rng.randn(M).astype(np.float32)
    return obs, act, old_logp, adv, ret


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, oracle.LIAR3_SPACE])
@pytest.mark.parametrize("ent_coef", [0.0, 0.01])
def test_single_minibatch_gradient_and_step_match_torch_autograd(kw, ent_coef):
    M = 300  # 3 tiles, the last one ragged
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=10)
    flat0 = pol.to_flat()
    space = oracle.make_space(**kw)
    assert flat0.size == oracle.param_count(space)
    obs, act, old_logp, adv, ret = make_batch(kw, M)
    nslot, nh = len(kw["nvec"]), len(kw["heads"])
    # move the policy off its init so ratios leave the clip range for some samples
    with torch.no_grad():
        for p in pol.ordered_parameters():
            p.add_(0.05 * torch.randn_like(p))
    flat0 = pol.to_flat()
    ev = oracle.policy_forward(space, flat0, obs, action_in=act)
    lg, val = pol.forward_logits(obs[:, :nslot])
    assert np.allclose(ev["logits"], lg.numpy(), atol=2e-6) and np.allclose(ev["value"], val.numpy(), atol=2e-6)
    old_logp = (ev["logp"] + 0.3 * np.random.RandomState(1).randn(M)).astype(np.float32)
    ref = sb3_torch.ppo_minibatch_step(pol, obs[:, :nslot], act[:, :nh], old_logp, adv, ret, ent_coef=ent_coef)
    params = flat0.copy()
    m, v = np.zeros_like(params), np.zeros_like(params)
    for grid in (1, 2, 5):
        p_, m_, v_ = params.copy(), m.copy(), v.copy()
        stats, grad = oupd.ppo_update(space, p_, m_, v_, 0, obs, act, old_logp, adv, ret,
                                      np.arange(M, dtype=np.int32)[None], M, grid, ent_coef=ent_coef)
        gscale = max(1.0, np.abs(ref["grad"]).max())
        assert np.abs(grad - ref["grad"]).max() <= 2e-6 * gscale, np.abs(grad - ref["grad"]).max()
        assert np.allclose(p_, pol.to_flat(), atol=3e-7, rtol=0)  # one Adam step of lr 3e-4
        st = stats[0]
        for i, k in enumerate(("pg_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction", "loss", "grad_norm")):
            assert abs(st[i] - ref[k]) <= 2e-5 * max(1.0, abs(ref[k])), (k, st[i], ref[k])
        assert st[7] == M
        assert 0.05 < st[4] < 0.95  # the clip branch is exercised both ways


def test_full_train_matches_torch_over_epochs():
    kw = oracle.LIAR_SPACE
    M, BS, E = 700, 256, 3   # 3 minibatches per epoch, the last short (188)
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=3)
    space = oracle.make_space(**kw)
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=2)
    ev = oracle.policy_forward(space, pol.to_flat(), obs, action_in=act)
    old_logp = (ev["logp"] + 0.05 * np.random.RandomState(5).randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    params = pol.to_flat().copy()
    m, v = np.zeros_like(params), np.zeros_like(params)
    stats, _ = oupd.ppo_update(space, params, m, v, 0, obs, act, old_logp, adv, ret, perm, BS, grid=3)
    ref = sb3_torch.ppo_train(pol, obs[:, :30], act[:, :2], old_logp, adv, ret, perm, BS)
    assert len(ref) == stats.shape[0] == 9
    assert np.allclose(params, pol.to_flat(), atol=5e-6, rtol=0)
    for i, k in enumerate(("pg_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction", "loss")):
        want = np.array([r[k] for r in ref])
        assert np.allclose(stats[:, i], want, atol=5e-5, rtol=1e-4), k
    assert np.array_equal(stats[:, 7], [256, 256, 188] * 3)
    # Adam state continues across calls: a second train() from step 9
    stats2, _ = oupd.ppo_update(space, params, m, v, 9, obs, act, old_logp, adv, ret, perm[:1], BS, grid=3)
    ref2 = sb3_torch.ppo_train(pol, obs[:, :30], act[:, :2], old_logp, adv, ret, perm[:1], BS)
    assert np.allclose(params, pol.to_flat(), atol=1e-5, rtol=0)


def test_normalize_advantage_guard_and_index_indirection():
    kw = oracle.RPS_SPACE
    space = oracle.make_space(**kw)
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=1)
    M = 64
    obs, act, old_logp, adv, ret = make_batch(kw, 3 * M, seed=7)
    index = (np.arange(M) * 3 + 1).astype(np.int32)   # every third sample
    params = pol.to_flat().copy()
    m, v = np.zeros_like(params), np.zeros_like(params)
    perm = oupd.perm_feistel(M, 1, 1, 4)
    oupd.ppo_update(space, params, m, v, 0, obs, act, old_logp, adv, ret, perm, 16, grid=2, index=index,
                    normalize_advantage=False)
    ref = sb3_torch.ppo_train(pol, obs[index][:, :1], act[index][:, :1], old_logp[index], adv[index],
                              ret[index], perm, 16, normalize_advantage=False)
    assert np.allclose(params, pol.to_flat(), atol=2e-6, rtol=0)


@pytest.mark.parametrize("M", [1, 2, 5, 64, 1000, 4097, 70000])
def test_feistel_shuffle_is_a_permutation(M):
    perm = oupd.perm_feistel(M, 3, seed=10, stream=4, epoch0=7)
    for e in range(3):
        assert np.array_equal(np.sort(perm[e]), np.arange(M))
    if M >= 64:
        assert not np.array_equal(perm[0], perm[1])
        assert not np.array_equal(perm[0], np.arange(M))
        # crude uniformity: displacement has no strong structure
        assert abs(np.corrcoef(perm[0], np.arange(M))[0, 1]) < 0.2
    again = oupd.perm_feistel(M, 1, seed=10, stream=4, epoch0=8)
    assert np.array_equal(again[0], perm[1])


def test_index_build_is_env_major():
    count = np.array([2, 0, 3, 1], np.int32)
    idx = oupd.index_build(count, T=3, N=4)
    assert idx.tolist() == [0, 4, 2, 6, 10, 3]
    dense = oupd.index_build(None, T=2, N=3)
    assert dense.tolist() == [0, 3, 1, 4, 2, 5]   # SB3 swap_and_flatten


def test_oracle_bc_matches_torch_autograd():
    """loss_kind = 1 (behaviour cloning, pantheonrl/algos/bc.py:270-315 + torch Adam defaults) against an
    independent torch-autograd restatement: parameters after 2 epochs x 5 batches and the logged stats."""
    kw = oracle.LIAR_SPACE
    space = oracle.make_space(**kw)
    M, BS, E = 150, 32, 2
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=9)
    params = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=2)
    perm = oupd.perm_feistel(M, E, seed=4, stream=5)
    m, v = np.zeros_like(params), np.zeros_like(params)
    hp = dict(loss_kind=1, l2_weight=0.01, ent_coef=1e-3, vf_coef=0.0, max_grad_norm=float("inf"),
              learning_rate=1e-3, eps=1e-8, normalize_advantage=False)
    stats, _ = oupd.ppo_update(space, params, m, v, 0, obs, act, old_logp, adv, ret, perm, BS, grid=2, **hp)
    ref = sb3_torch.bc_train(pol, obs[:, :30], act[:, :2], perm, BS, ent_weight=1e-3, l2_weight=0.01)
    assert np.allclose(params, pol.to_flat(), atol=2e-5, rtol=0)
    assert np.allclose(stats[:, 0], [r["neglogp"] for r in ref], atol=1e-5)
    assert np.allclose(-stats[:, 2], [r["entropy"] for r in ref], atol=1e-5)
    assert np.allclose(stats[:, 3], [r["prob_true_act"] for r in ref], atol=1e-6)
    assert stats[-1, 0] < stats[0, 0]  # it learns: neglogp of the cloned actions goes down


def test_uniform_slot_rows_equal_the_bias_gradient_bit_for_bit():
    """Property behind a round-2 kernel lead: when every sample of a tile holds the same value in
    a one-hot slot, the first-layer weight-gradient row that value selects is the same ascending chain over the
    tile's samples as the first-layer bias gradient — in both towers, exactly."""
    kw = oracle.LIAR_SPACE
    space = oracle.make_space(**kw)
    M = 128  # one tile, one minibatch, one CTA
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=3)
    params = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=8)
    uniform = {7: 6, 8: 0, 20: 3, 29: 11}  # slot -> the value all samples hold
    for s, val in uniform.items():
        obs[:, s] = val
    perm = np.arange(M, dtype=np.int32)[None]
    m, v = np.zeros_like(params), np.zeros_like(params)
    _, grad = oupd.ppo_update(space, params, m, v, 0, obs, act, old_logp, adv, ret, perm, M, grid=1, ent_coef=0.01)
    F, H = sum(kw["nvec"]), 64
    off = np.concatenate([[0], np.cumsum(kw["nvec"])])
    for base_w, base_b in ((0, F * H), (F * H + H + H * H + H, F * H + H + H * H + H + F * H)):  # pi tower, vf tower
        bias = grad[base_b:base_b + H]
        assert np.any(bias != 0)
        for s, val in uniform.items():
            row = off[s] + val
            assert np.array_equal(grad[base_w + row * H: base_w + (row + 1) * H], bias), (s, val)
            for other in range(kw["nvec"][s]):
                if other != val:
                    r = off[s] + other
                    assert not grad[base_w + r * H: base_w + (r + 1) * H].any()
