"""GPU parity of the ModularAlgorithm path (pantheonrl/algos/modular): one partner phase of ModularAlgorithm.train
per launch of the update kernel (loss_kind PTH_LOSS_MODULAR) vs the CPU oracle, bit for bit, and against the
parameters the reference's own ModularAlgorithm.train produced (tests/golden/modular.npz, tolerance)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import update as oupd
from pantheonrl_b200 import _lib, update as dupd
from test_oracle_update import make_batch

pytestmark = pytest.mark.gpu
BOX = dict(box_dim=62, heads=[6])


def d(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def gspace(kw):
    return _lib.Space.box(kw["box_dim"], kw["heads"]) if "box_dim" in kw else _lib.Space.onehot(kw["nvec"], kw["heads"])


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


def run_phase(kw, Pn, q, gp, gm, gv, op, om, ov, step, vf_step, obs, act, old_logp, adv, ret, perm, BS, grid, coef):
    """One partner phase on the GPU (in place on the device tensors) and on the oracle (in place on the arrays)."""
    osp, sp = oracle.make_space(**kw), gspace(kw)
    M = perm.shape[1]
    n = perm.shape[0] * (-(-M // BS))
    ws = dupd.UpdateWorkspace(sp, M, BS, num_partners=Pn)
    mg = torch.full((n,), -1.0, device="cuda")
    gst = dupd.ppo_update(sp, gp, gm, gv, step, d(obs), d(act), d(old_logp), d(adv), d(ret), d(perm), BS, ws,
                          grid_ctas=grid, ent_coef=0.01, loss_kind=_lib.PTH_LOSS_MODULAR, num_partners=Pn, partner_idx=q,
                          partner_vf_step=vf_step, marginal_reg_coef=coef, ctx_loss=mg)
    torch.cuda.synchronize()
    ost, omg = oupd.modular_update(osp, op, om, ov, step, vf_step, Pn, q, obs, act, old_logp, adv, ret, perm, BS, grid,
                                   ent_coef=0.01, marginal_reg_coef=coef)
    gst = gst.cpu().numpy()
    assert np.array_equal(gst, ost), (q, np.abs(gst - ost).max())
    assert np.array_equal(mg.cpu().numpy(), omg), q
    assert np.array_equal(gm.cpu().numpy(), om) and np.array_equal(gv.cpu().numpy(), ov), q
    assert np.array_equal(gp.cpu().numpy(), op), (q, np.abs(gp.cpu().numpy() - op).max())
    return gst, omg


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, BOX])
@pytest.mark.parametrize("Pn,M,BS,E,grid,coef", [(1, 300, 300, 1, 2, 0.0), (2, 280, 64, 2, 1, 0.5), (3, 700, 256, 2, 3, 1.0),
                                                (8, 600, 300, 1, 96, 0.3)])
def test_modular_train_bit_exact_vs_oracle(ctx, kw, Pn, M, BS, E, grid, coef):
    """Two rounds of partner phases (the second with non-zero Adam steps for every group): several tiles per CTA,
    more CTAs than tiles, a ragged last tile, 1 to 8 partner modules."""
    osp = oracle.make_space(**kw)
    rs = np.random.RandomState(M + Pn)
    P = oupd.modular_param_count(osp, Pn)
    params = (0.25 * rs.randn(P)).astype(np.float32)
    gp, gm, gv = d(params), d(np.zeros(P, np.float32)), d(np.zeros(P, np.float32))
    op, om, ov = params.copy(), np.zeros(P, np.float32), np.zeros(P, np.float32)
    step, vf_steps = 0, [0] * Pn
    for rnd in range(2):
        for q in range(Pn if Pn <= 3 else 2):
            obs, act, adv, ret = batch(kw, M, M + 7 * q + rnd)
            ev = oupd.modular_forward(osp, op, Pn, q, obs, action_in=act)
            old_logp = (ev["logp"] + 0.1 * rs.randn(M)).astype(np.float32)
            perm = oupd.perm_feistel(M, E, seed=10 + rnd, stream=4 + q)
            st, mg = run_phase(kw, Pn, q, gp, gm, gv, op, om, ov, step, vf_steps[q], obs, act, old_logp, adv, ret, perm,
                               BS, grid, coef)
            step += st.shape[0]
            vf_steps[q] += st.shape[0]
            assert np.all(np.isfinite(st)) and np.all(mg >= 0)


@pytest.mark.parametrize("name,kw", [("rps2", oracle.RPS_SPACE), ("liar3", oracle.LIAR_SPACE), ("liar1", oracle.LIAR_SPACE)])
def test_modular_train_reproduces_the_reference_run(ctx, name, kw):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "modular.npz"))
    pre = name + "_"
    Pn, BS, E = (int(x) for x in g[pre + "hp"])
    coef = float(g[pre + "coef"][0])
    log = dict(zip(g[pre + "log_keys"], g[pre + "log_vals"]))
    p0 = g[pre + "p0"]
    gp, gm, gv = d(p0), d(np.zeros_like(p0)), d(np.zeros_like(p0))
    op, om, ov = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    step, sts = 0, []
    for q in range(Pn):
        b = {k: g[f"{pre}b{q}_{k}"] for k in ("obs", "act", "old_logp", "adv", "ret", "perms")}
        st, _ = run_phase(kw, Pn, q, gp, gm, gv, op, om, ov, step, 0, b["obs"], b["act"], b["old_logp"], b["adv"], b["ret"],
                          b["perms"], BS, 3, coef)
        step += st.shape[0]
        sts.append(st)
    st = np.concatenate(sts)
    assert np.abs(gp.cpu().numpy() - g[pre + "params"]).max() <= 2e-6
    assert st[:, 0].mean() == pytest.approx(log["train/policy_gradient_loss"], abs=2e-6)
    assert st[:, 1].mean() == pytest.approx(log["train/value_loss"], abs=2e-5)
    assert st[:, 2].mean() == pytest.approx(log["train/entropy_loss"], abs=2e-5)


@pytest.mark.parametrize("kw", [oracle.RPS_SPACE, oracle.LIAR_SPACE, BOX])
@pytest.mark.parametrize("Pn,q,B", [(1, 0, 1), (3, 2, 777), (8, 5, 129)])
def test_modular_forward_bit_exact(ctx, kw, Pn, q, B):
    """ModularPolicy.forward / evaluate_actions: logits = main + partner[q], value = main + partner[q]."""
    from pantheonrl_b200 import ops
    osp, sp = oracle.make_space(**kw), gspace(kw)
    params = (0.3 * np.random.RandomState(B + Pn).randn(oupd.modular_param_count(osp, Pn))).astype(np.float32)
    obs, act, _, _ = batch(kw, B, B + 1)
    for action_in in (None, act):
        want = oupd.modular_forward(osp, params, Pn, q, obs, seed=3, tick=9, idx0=5, action_in=action_in)
        got = ops.policy_forward(sp, d(params), d(obs), seed=3, tick=9, idx0=5, num_partners=Pn, partner_idx=q,
                                 action_in=None if action_in is None else d(action_in))
        for k in ("action", "value", "logp", "entropy", "logits"):
            assert np.array_equal(got[k].cpu().numpy(), want[k]), (k, action_in is None)
