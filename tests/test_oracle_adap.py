"""The ADAP half of the oracle against the reference's OWN code.

tests/golden/adap.npz was produced by executing, verbatim, `ADAP.train` (pantheonrl/algos/adap/adap_learn.py:229-347)
with its context loss ON — `get_context_kl_loss` / `kl_divergence` / `SAMPLERS` of pantheonrl/algos/adap/util.py and
`AdapPolicy._get_latent` / `evaluate_actions` of pantheonrl/algos/adap/policies.py — and recording the random draws the
context loss made (tests/golden/make_golden_adap.py).  Here the same inputs and draws go through our restatements:
oracle/sb3_torch.py (torch eager + autograd) and the C oracle's `orc_ppo_update` with `loss_kind = 2` (hand-written
backward of the context loss; the thing the CUDA kernel is bit-exact with)."""
import os

import numpy as np
import pytest

import oracle
from oracle import sb3_torch
from oracle import update as oupd

CASES = [("rps", oracle.RPS_SPACE), ("liar", oracle.LIAR_SPACE), ("liar_k3", oracle.LIAR_SPACE)]
MULT_CASES = [("mult_rps", oracle.RPS_SPACE), ("mult_liar", oracle.LIAR_SPACE)]  # AdapPolicyMult (ADAP_MULT)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "adap.npz"))


def _case(g, name):
    pre = name + "_"
    d = {k[len(pre):]: g[k] for k in g.files if k.startswith(pre) and not (name == "liar" and k.startswith("liar_k3_"))}
    assert "hp" in d, name
    d["log"] = dict(zip(d["log_keys"], d["log_vals"]))
    return d


@pytest.mark.parametrize("name,kw", CASES)
def test_adap_train_matches_the_reference(g, name, kw):
    d = _case(g, name)
    M, BS, E, K, S = (int(x) for x in d["hp"])
    coeff = float(d["coeff"][0])
    nslot, nh = len(kw["nvec"]), len(kw["heads"])
    n_mb = -(-M // BS)
    log, want = d["log"], d["params"]
    assert d["sidx"].shape == (E * n_mb, S) and d["draws"].shape == (E * n_mb, K, 3)

    # 1. torch-eager restatement (autograd through the context loss)
    pol = sb3_torch.AdapMlpPolicy(nvec=kw["nvec"], heads=kw["heads"], context_size=3, seed=0)
    pol.from_flat(d["p0"])
    full_obs = np.concatenate([d["obs"][:, :nslot].astype(np.float32), d["ctx"]], axis=1)
    stats = sb3_torch.adap_train(pol, full_obs, d["act"][:, :nh], d["old_logp"], d["adv"], d["ret"], d["perms"], BS,
                                 d["sidx"], d["draws"], context_loss_coeff=coeff, ent_coef=0.01)
    assert np.abs(pol.to_flat() - want).max() <= 1e-6  # same op sequence; torch's CPU GEMM blocking varies with threads / load
    assert np.mean([s["context_loss"] for s in stats[-n_mb:]]) == pytest.approx(log["train/context_kl_loss"], abs=1e-6)
    assert stats[-1]["loss"] == pytest.approx(log["train/loss"], abs=1e-6)

    # 2. the C oracle with one and with three lane groups (context tiles join the tile -> CTA round robin)
    space = oracle.make_space(**kw)
    for grid in (1, 3):
        p, m, v = d["p0"].copy(), np.zeros_like(d["p0"]), np.zeros_like(d["p0"])
        st, _, cl = oupd.ppo_update(space, p, m, v, 0, d["obs"], d["act"], d["old_logp"], d["adv"], d["ret"], d["perms"],
                                    BS, grid=grid, ent_coef=0.01, loss_kind=2, ctx=d["ctx"], ctx_loss_coeff=coeff,
                                    ctx_sidx=d["sidx"], ctx_draws=d["draws"])
        assert np.abs(p - want).max() <= 5e-6, grid
        assert cl[-n_mb:].mean() == pytest.approx(log["train/context_kl_loss"], abs=1e-5)
        assert st[-1, 5] == pytest.approx(log["train/loss"], abs=2e-5)
        assert st[:, 0].mean() == pytest.approx(log["train/policy_gradient_loss"], abs=2e-6)
        assert st[:, 1].mean() == pytest.approx(log["train/value_loss"], abs=2e-5)


def test_context_loss_moves_the_parameters(g):
    """The pin above would be hollow if the context term were too small to matter: without it the oracle's
    parameters differ from the reference's by far more than the tolerance."""
    d = _case(g, "liar_k3")
    M, BS, E, K, S = (int(x) for x in d["hp"])
    space = oracle.make_space(**oracle.LIAR_SPACE)
    p, m, v = d["p0"].copy(), np.zeros_like(d["p0"]), np.zeros_like(d["p0"])
    oupd.ppo_update(space, p, m, v, 0, d["obs"], d["act"], d["old_logp"], d["adv"], d["ret"], d["perms"], BS, grid=1,
                    ent_coef=0.01, loss_kind=0, ctx=d["ctx"])
    assert np.abs(p - d["params"]).max() > 1e-4


def test_adap_policy_without_context_columns_is_the_mlp_policy():
    """context_size inputs whose weights are zero change nothing: AdapPolicy's forward equals MlpPolicy's."""
    rs = np.random.RandomState(0)
    kw = oracle.LIAR_SPACE
    space = oracle.make_space(**kw)
    P0, F = oracle.param_count(space), sum(kw["nvec"])
    base = (0.1 * rs.randn(P0)).astype(np.float32)
    obs = np.zeros((50, 32), np.uint8)
    for s_, n in enumerate(kw["nvec"]):
        obs[:, s_] = rs.randint(0, n, 50)
    # widen both first-layer matrices by 3 zero rows
    parts, o = [], 0
    for n in (64 * F, 64, 64 * 64, 64, 64 * F, 64, 64 * 64, 64, P0 - 2 * (64 * F + 64 + 64 * 64 + 64)):
        parts.append(base[o:o + n])
        o += n
    wide = np.concatenate([parts[0], np.zeros(3 * 64, np.float32), *parts[1:4], parts[4], np.zeros(3 * 64, np.float32),
                           *parts[5:]])
    assert wide.size == oracle.adap_param_count(space, 3)
    a = oracle.policy_forward(space, base, obs, seed=5)
    b = oracle.adap_forward(space, wide, obs, rs.randn(50, 3).astype(np.float32), seed=5)
    for k in ("action", "value", "logp", "entropy", "logits"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name,kw", MULT_CASES)
def test_adap_mult_train_matches_the_reference(g, name, kw):
    """AdapPolicyMult: the golden run used the reference's own MultModel (adap/policies.py:134-267) as the
    mlp_extractor under AdapPolicy's methods and ADAP.train with the context loss on."""
    d = _case(g, name)
    M, BS, E, K, S = (int(x) for x in d["hp"])
    coeff = float(d["coeff"][0])
    nslot, nh = len(kw["nvec"]), len(kw["heads"])
    n_mb = -(-M // BS)
    log, want = d["log"], d["params"]
    pol = sb3_torch.AdapMultPolicy(nvec=kw["nvec"], heads=kw["heads"], context_size=3, seed=0)
    pol.from_flat(d["p0"])
    full_obs = np.concatenate([d["obs"][:, :nslot].astype(np.float32), d["ctx"]], axis=1)
    stats = sb3_torch.adap_train(pol, full_obs, d["act"][:, :nh], d["old_logp"], d["adv"], d["ret"], d["perms"], BS,
                                 d["sidx"], d["draws"], context_loss_coeff=coeff, ent_coef=0.01)
    assert np.abs(pol.to_flat() - want).max() <= 1e-6
    assert np.mean([s_["context_loss"] for s_ in stats[-n_mb:]]) == pytest.approx(log["train/context_kl_loss"], abs=1e-6)
    space = oracle.make_space(**kw)
    assert want.size == oracle.adap_mult_param_count(space, 3)
    for grid in (1, 3):
        p, m, v = d["p0"].copy(), np.zeros_like(d["p0"]), np.zeros_like(d["p0"])
        st, _, cl = oupd.ppo_update(space, p, m, v, 0, d["obs"], d["act"], d["old_logp"], d["adv"], d["ret"], d["perms"],
                                    BS, grid=grid, ent_coef=0.01, loss_kind=2, ctx=d["ctx"], ctx_loss_coeff=coeff,
                                    ctx_sidx=d["sidx"], ctx_draws=d["draws"], adap_mult=True)
        assert np.abs(p - want).max() <= 2e-6, grid
        assert cl[-n_mb:].mean() == pytest.approx(log["train/context_kl_loss"], abs=1e-5)
        assert st[-1, 5] == pytest.approx(log["train/loss"], abs=2e-5)
    # forward: the oracle's AdapPolicyMult equals the torch modules
    import torch as th
    ev = oracle.adap_mult_forward(space, want, d["obs"][:64], d["ctx"][:64], action_in=d["act"][:64])
    pol.from_flat(want)
    with th.no_grad():
        values, logp, ent = pol.evaluate_actions(full_obs[:64], d["act"][:64, :nh])
    assert np.abs(ev["value"] - values.numpy().reshape(-1)).max() < 1e-5 and np.abs(ev["logp"] - logp.numpy()).max() < 1e-5
