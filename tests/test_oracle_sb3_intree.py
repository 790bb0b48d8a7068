"""The arithmetic half of the oracle against the reference's OWN in-tree code.

tests/golden/sb3_intree.npz was produced by executing, verbatim,
  * `ADAP.train` (pantheonrl/algos/adap/adap_learn.py:229-347) — the reference's copy of SB3's
    `PPO.train` — with the extra context loss switched off, and
  * the GAE bootstrap loop of overcookedgym/.../baselines/ppo2/runner.py:152-164
(tests/golden/make_golden_sb3_intree.py).  Here the same inputs go through our restatements:
oracle/sb3_torch.py (torch eager), the C oracle's `orc_ppo_update` (the thing the CUDA kernel is
bit-exact with), oracle/sb3_numpy.py and `orc_gae`."""
import os

import numpy as np
import pytest

import oracle
from oracle import sb3_numpy, sb3_torch
from oracle import update as oupd


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "sb3_intree.npz"))


CASES = [("rps", oracle.RPS_SPACE), ("liar", oracle.LIAR_SPACE)]


@pytest.mark.parametrize("name,kw", CASES)
def test_ppo_train_matches_the_references_in_tree_copy(g, name, kw):
    pre = f"train_{name}_"
    M, BS, E = (int(x) for x in g[pre + "hp"])
    obs, act, perms = g[pre + "obs"], g[pre + "act"], g[pre + "perms"]
    old_logp, adv, ret, p0, want = g[pre + "old_logp"], g[pre + "adv"], g[pre + "ret"], g[pre + "p0"], g[pre + "params"]
    log = dict(zip(g[pre + "log_keys"], g[pre + "log_vals"]))
    nslot, nh = len(kw["nvec"]), len(kw["heads"])
    n_mb = -(-M // BS)

    # 1. torch-eager restatement: same ops in the same order -> (practically) the same bits
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=0)
    pol.from_flat(p0)
    stats = sb3_torch.ppo_train(pol, obs[:, :nslot], act[:, :nh], old_logp, adv, ret, perms, BS, ent_coef=0.01)
    assert np.abs(pol.to_flat() - want).max() <= 1e-7
    assert np.mean([s["pg_loss"] for s in stats]) == pytest.approx(log["train/policy_gradient_loss"], abs=1e-7)
    assert np.mean([s["value_loss"] for s in stats]) == pytest.approx(log["train/value_loss"], abs=1e-6)
    assert np.mean([s["entropy_loss"] for s in stats]) == pytest.approx(log["train/entropy_loss"], abs=1e-6)
    assert np.mean([s["clip_fraction"] for s in stats]) == pytest.approx(log["train/clip_fraction"], abs=1e-7)
    assert stats[-1]["loss"] == pytest.approx(log["train/loss"], abs=1e-6)
    # SB3 resets approx_kl_divs every epoch: the logged value is the LAST epoch's mean
    assert np.mean([s["approx_kl"] for s in stats[-n_mb:]]) == pytest.approx(log["train/approx_kl"], abs=1e-7)
    assert int(g[pre + "n_updates"]) == E

    # 2. the C oracle (what the CUDA kernel equals bit for bit): fp32, its own summation order
    space = oracle.make_space(**kw)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    st, _ = oupd.ppo_update(space, p, m, v, 0, obs, act, old_logp, adv, ret, perms, BS, grid=3, ent_coef=0.01)
    assert np.abs(p - want).max() <= 5e-6
    assert st[:, 0].mean() == pytest.approx(log["train/policy_gradient_loss"], abs=2e-6)
    assert st[:, 1].mean() == pytest.approx(log["train/value_loss"], abs=2e-5)
    assert st[:, 2].mean() == pytest.approx(log["train/entropy_loss"], abs=2e-5)
    assert st[-n_mb:, 3].mean() == pytest.approx(log["train/approx_kl"], abs=2e-6)
    assert st[:, 4].mean() == pytest.approx(log["train/clip_fraction"], abs=1e-3)  # a ratio within 1 ulp of the clip edge may flip
    assert st[-1, 5] == pytest.approx(log["train/loss"], abs=2e-5)


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_gae_matches_the_in_tree_bootstrap_loop(g, case):
    pre = f"gae_{case}_"
    rew, val, start = g[pre + "rewards"], g[pre + "values"], g[pre + "episode_starts"]
    lv, dn, adv_ref, ret_ref = g[pre + "last_values"], g[pre + "dones"], g[pre + "advantages"], g[pre + "returns"]
    # the in-tree loop accumulates in float64 (1.0 - bool array) and stores float32; ours is fp32 throughout:
    # the north star's tolerance for returns is 1e-5
    adv, ret = oracle.gae(rew, val, start, lv, dn)
    assert np.abs(adv - adv_ref).max() <= 1e-5 and np.abs(ret - ret_ref).max() <= 1e-5
    adv2, ret2 = sb3_numpy.compute_returns_and_advantage(rew, val, start, lv, dn)
    assert np.array_equal(adv2, adv) and np.array_equal(ret2, ret)


@pytest.mark.parametrize("name,kw", CASES)
def test_behaviour_cloning_matches_the_references_bc_train(g, name, kw):
    """`BC.train` + `BC._calculate_loss` of pantheonrl/algos/bc.py:270-357, executed verbatim by the golden
    generator, vs the torch restatement and the C oracle's loss_kind = 1 (what PTH_LOSS_BC equals bit for bit)."""
    pre = f"bc_{name}_"
    M, BS, E = (int(x) for x in g[pre + "hp"])
    ent_w, l2 = (float(x) for x in g[pre + "w"])
    obs, act, perms, p0, want, ref = g[pre + "obs"], g[pre + "act"], g[pre + "perms"], g[pre + "p0"], g[pre + "params"], g[pre + "stats"]
    nslot, nh = len(kw["nvec"]), len(kw["heads"])
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=0)
    pol.from_flat(p0)
    st = sb3_torch.bc_train(pol, obs[:, :nslot], act[:, :nh], perms, BS, ent_weight=ent_w, l2_weight=l2)
    assert np.abs(pol.to_flat() - want).max() <= 1e-7
    assert np.allclose([s["neglogp"] for s in st], ref[:, 0], atol=1e-6) and np.allclose([s["loss"] for s in st], ref[:, 3], atol=1e-5)
    space = oracle.make_space(**kw)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    z = np.zeros(M, np.float32)
    cst, _ = oupd.ppo_update(space, p, m, v, 0, obs, act, z, z, z, perms, BS, grid=2, loss_kind=1, l2_weight=l2,
                             ent_coef=ent_w, vf_coef=0.0, max_grad_norm=float("inf"), learning_rate=1e-3, eps=1e-8,
                             normalize_advantage=False)
    assert np.abs(p - want).max() <= 2e-5  # Adam at eps 1e-8 amplifies fp32 summation-order differences of tiny gradients
    assert np.allclose(cst[:, 0], ref[:, 0], atol=2e-5)          # neglogp
    assert np.allclose(-cst[:, 2], ref[:, 1], atol=2e-5)         # entropy
    assert np.allclose(cst[:, 3], ref[:, 2], atol=1e-6)          # prob_true_act
