"""The ModularAlgorithm half of the oracle against the reference's OWN code.

tests/golden/modular.npz was produced by executing, verbatim, `ModularAlgorithm.train`
(pantheonrl/algos/modular/learn.py:221-351) over per-partner buffers with the methods of `ModularPolicy`
(pantheonrl/algos/modular/policies.py:273-396) bound to the torch modules of oracle/sb3_torch.ModularMlpPolicy
(tests/golden/make_golden_modular.py).  Here the same inputs go through oracle/sb3_torch.modular_train (torch eager +
autograd) and the C oracle's `orc_modular_update` (hand-written backward through the partner modules and the
marginal regulariser; the thing the CUDA kernel is bit-exact with), one call per partner phase."""
import os

import numpy as np
import pytest

import oracle
from oracle import sb3_torch
from oracle import update as oupd

CASES = [("rps2", oracle.RPS_SPACE), ("liar3", oracle.LIAR_SPACE), ("liar1", oracle.LIAR_SPACE)]


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "modular.npz"))


def phases(g, name):
    pre = name + "_"
    Pn, BS, E = (int(x) for x in g[pre + "hp"])
    bufs = [{k: g[f"{pre}b{p}_{k}"] for k in ("obs", "act", "old_logp", "adv", "ret", "perms")} for p in range(Pn)]
    return Pn, BS, E, float(g[pre + "coef"][0]), bufs, dict(zip(g[pre + "log_keys"], g[pre + "log_vals"]))


def run_oracle(space, p, Pn, BS, E, coef, bufs, grid):
    m, v = np.zeros_like(p), np.zeros_like(p)
    step, sts, mgs = 0, [], []
    for q, b in enumerate(bufs):
        st, mg = oupd.modular_update(space, p, m, v, step, 0, Pn, q, b["obs"], b["act"], b["old_logp"], b["adv"],
                                     b["ret"], b["perms"], BS, grid, ent_coef=0.01, marginal_reg_coef=coef)
        step += st.shape[0]
        sts.append(st)
        mgs.append(mg)
    return m, v, np.concatenate(sts), np.concatenate(mgs)


@pytest.mark.parametrize("name,kw", CASES)
def test_modular_train_matches_the_reference(g, name, kw):
    Pn, BS, E, coef, bufs, log = phases(g, name)
    nslot, nh = len(kw["nvec"]), len(kw["heads"])
    want = g[name + "_params"]
    pol = sb3_torch.ModularMlpPolicy(nvec=kw["nvec"], heads=kw["heads"], num_partners=Pn, seed=0)
    pol.from_flat(g[name + "_p0"])
    stats = sb3_torch.modular_train(pol, [(b["obs"][:, :nslot], b["act"][:, :nh], b["old_logp"], b["adv"], b["ret"])
                                          for b in bufs], [b["perms"] for b in bufs], BS, marginal_reg_coef=coef,
                                    ent_coef=0.01)
    assert np.abs(pol.to_flat() - want).max() <= 1e-6  # same op sequence; torch's CPU GEMM blocking may differ with load
    space = oracle.make_space(**kw)
    assert want.size == oupd.modular_param_count(space, Pn)
    for grid in (1, 3):
        p = g[name + "_p0"].copy()
        _, _, st, mg = run_oracle(space, p, Pn, BS, E, coef, bufs, grid)
        assert np.abs(p - want).max() <= 2e-6, grid
        assert st[:, 0].mean() == pytest.approx(log["train/policy_gradient_loss"], abs=2e-6)
        assert st[:, 1].mean() == pytest.approx(log["train/value_loss"], abs=2e-5)
        assert st[:, 2].mean() == pytest.approx(log["train/entropy_loss"], abs=2e-5)
        assert mg.mean() == pytest.approx(np.mean([s["marginal"] for s in stats]), abs=1e-6)


def test_marginal_regulariser_and_idle_value_modules(g):
    """Without the regulariser the result is far from the reference's (the pin is not hollow); the value modules of a
    partner are only touched in that partner's own phase (no gradient -> Adam skips them)."""
    Pn, BS, E, coef, bufs, _ = phases(g, "liar3")
    space = oracle.make_space(**oracle.LIAR_SPACE)
    p = g["liar3_p0"].copy()
    run_oracle(space, p, Pn, BS, E, 0.0, bufs, 1)
    assert np.abs(p - g["liar3_params"]).max() > 1e-4
    # one phase only (partner 1): the value modules of partners 0 and 2 keep their initial values and zero moments
    p, m, v = g["liar3_p0"].copy(), np.zeros_like(p), np.zeros_like(p)
    b = bufs[1]
    oupd.modular_update(space, p, m, v, 0, 0, Pn, 1, b["obs"], b["act"], b["old_logp"], b["adv"], b["ret"], b["perms"], BS,
                        2, ent_coef=0.01, marginal_reg_coef=coef)
    main = oracle.param_count(space)
    blk = (p.size - main) // Pn
    L = 19
    vf = slice(2 * (64 * 64 + 64), 4 * (64 * 64 + 64))
    val = slice(4 * (64 * 64 + 64) + L * 64 + L, blk)
    for q in range(Pn):
        base = main + q * blk
        for sl in (vf, val):
            same = np.array_equal(p[base:base + blk][sl], g["liar3_p0"][base:base + blk][sl])
            assert same == (q != 1), (q, sl)
            assert (not m[base:base + blk][sl].any()) == (q != 1)
        # the policy side of every partner moves (marginal regulariser)
        assert not np.array_equal(p[base:base + 64 * 64], g["liar3_p0"][base:base + 64 * 64])


def test_modular_forward_composes_main_and_partner():
    kw = oracle.LIAR_SPACE
    space = oracle.make_space(**kw)
    pol = sb3_torch.ModularMlpPolicy(nvec=kw["nvec"], heads=kw["heads"], num_partners=2, seed=4)
    import torch as th
    with th.no_grad():
        pol.action_net.weight.mul_(50)
        pol.partner_action_net[1].weight.mul_(80)
    rs = np.random.RandomState(1)
    obs = np.zeros((40, 32), np.uint8)
    for s_, n in enumerate(kw["nvec"]):
        obs[:, s_] = rs.randint(0, n, 40)
    act = np.zeros((40, 4), np.uint8)
    act[:, 0], act[:, 1] = rs.randint(0, 7, 40), rs.randint(0, 12, 40)
    for q in range(2):
        got = oupd.modular_forward(space, pol.to_flat(), 2, q, obs, action_in=act)
        with th.no_grad():
            values, logp, ent = pol.evaluate_actions(obs[:, :30], act[:, :2], partner_idx=q)
        assert np.abs(got["value"] - values.numpy().reshape(-1)).max() < 1e-5
        assert np.abs(got["logp"] - logp.numpy()).max() < 1e-5 and np.abs(got["entropy"] - ent.numpy()).max() < 1e-5
