"""CPU tests: pin the oracle's Overcooked restatement (oracle/pth_oracle_overcooked.inc) on
  (1) the reference's OWN golden vectors: state_featurization.pickle (overcooked_test.py:169-173) and
      test_full_traj.json (overcooked_test.py:135-142), re-encoded by tests/golden/make_golden_overcooked.py;
  (2) traces recorded from the reference's OvercookedMultiEnv driven through MultiAgentEnv.step/reset
      and through multi_step with random joint actions on several layouts."""
import os

import numpy as np
import pytest

import oracle
from oracle import overcooked as oc
from oracle import rollout as orc


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def test_reference_state_featurization_pickle(golden_dir):
    g = _load(golden_dir, "oc_ref_featurization.npz")
    d = oc.named_layouts(golden_dir)["simple"]
    L = oc.make_layout(d["grid"], d["start"], d["cook_time"], d["num_items"], d["delivery_reward"], horizon=400)
    assert g["feats"].shape == (5, 400, 2, 62)
    for ep in range(5):
        feats, sparse, shaped, dones, _ = oc.replay(L, g["actions"][ep])
        assert np.array_equal(feats[:400], g["feats"][ep].astype(np.float32)), ep
        assert np.array_equal(sparse, g["sparse"][ep])
        assert dones[-1] == 1 and not dones[:-1].any()
    # shaped rewards of the plain mdp are zero-weighted there (rew_shaping_params None -> zeros)
    assert g["sparse"].sum() == 980


def test_reference_full_trajectory_json(golden_dir):
    g = _load(golden_dir, "oc_full_traj.npz")
    grid = [str(r) for r in g["grid"]]
    L = oc.make_layout(grid, g["start"], int(g["cook_time"]), int(g["num_items"]), int(g["delivery_reward"]),
                       horizon=100, order_list=list(g["order_list"]))
    _, sparse, _, dones, states = oc.replay(L, g["actions"], want_states=True)
    assert np.array_equal(states, g["states"])
    assert np.array_equal(sparse, g["rewards"])
    assert not dones.any()
    assert g["rewards"].sum() == 20 and (g["states"][:, 3] == 4).any()  # a delivery and a held tomato occur


@pytest.mark.parametrize("layout", ["simple", "random1", "corridor", "scenario2_s"])
def test_random_joint_actions_through_multi_step(golden_dir, layout):
    g = _load(golden_dir, f"oc_random_{layout}.npz")
    L = oc.multienv_layout(golden_dir, layout)
    feats, sparse, shaped, dones, states = oc.replay(L, g["actions"], want_states=True)
    assert np.array_equal(states, g["states"])
    assert np.array_equal(feats[:, 0], g["obs0"].astype(np.float32))
    assert np.array_equal(feats[:, 1], g["obs1"].astype(np.float32))
    assert np.array_equal(sparse + shaped, g["rewards"])
    assert np.array_equal(dones, g["dones"])


def _box_rows(a):
    out = np.zeros(a.shape[:-1] + (oc.ROW,), np.float32)
    out[..., :a.shape[-1]] = a
    return out


@pytest.mark.parametrize("fname", ["oc_routing_simple_e0.npz", "oc_routing_simple_e1.npz",
                                   "oc_routing_unident_s_e0.npz", "oc_routing_random0_e1.npz"])
def test_scripted_rollout_reproduces_multiagentenv_trace(golden_dir, fname):
    """OvercookedMultiEnv behind MultiAgentEnv.step/reset: what the ego (SB3 collect_rollouts) and the
    partner (OnPolicyAgent) see, episode boundaries included."""
    g = _load(golden_dir, fname)
    T = g["ego_act"].shape[0]
    L = oc.multienv_layout(golden_dir, str(g["layout"]), horizon=int(g["horizon"]))
    rows, latch = orc.partner_rows_from_events(dict(g, ev_pid=np.zeros(len(g["ev_kind"]), np.int32)))
    assert len(rows) == T
    pad4 = lambda a: np.pad(np.asarray(a, np.uint8).reshape(-1, 1), ((0, 0), (0, 3)))  # noqa: E731
    space = oracle.make_space(box_dim=62, heads=[6])
    ego, alt, carry = orc.rollout("overcooked", space, None, None, N=1, T=T,
                                  script_ego_act=pad4(g["ego_act"]),
                                  script_alt_act=pad4([r["act"] for r in rows]),
                                  oc_layout=L, oc_ego_idx=int(g["ego_idx"]))
    assert np.array_equal(ego["obs"][:, 0], _box_rows(g["ego_obs"]))
    assert np.array_equal(ego["rewards"][:, 0], g["ego_rew"])
    starts = np.concatenate([[1.0], g["ego_done"][:-1].astype(np.float32)])
    assert np.array_equal(ego["episode_starts"][:, 0], starts)
    assert carry["ego_last_done"][0] == float(g["ego_done"][-1])
    assert alt["count"][0] == T
    assert np.array_equal(alt["obs"][:T, 0], _box_rows(np.array([r["obs"] for r in rows])))
    assert np.array_equal(alt["rewards"][:T, 0], np.array([r["rew"] for r in rows], np.float32))
    assert np.array_equal(alt["episode_starts"][:T, 0], np.array([r["start"] for r in rows], np.float32))
    assert carry["alt_last_done"][0] == float(latch)
    assert g["ego_done"].sum() >= 1 and carry["ep_stats"][0] == g["ego_done"].sum()
    # the observation SB3 would bootstrap from
    feats = np.zeros((2, 62), np.float32)
    oracle.lib().orc_oc_featurize(__import__("ctypes").byref(L), carry["oc_state"].ctypes.data_as(
        __import__("ctypes").c_void_p), feats[0].ctypes.data_as(__import__("ctypes").c_void_p),
        feats[1].ctypes.data_as(__import__("ctypes").c_void_p))
    assert np.array_equal(feats[int(g["ego_idx"])], g["final_obs"].astype(np.float32))


def test_layout_tables_cover_every_trainer_layout(golden_dir):
    """Every layout trainer.py accepts (overcooked_utils.LAYOUT_LIST) fits the limits."""
    n = 0
    for name, d in oc.named_layouts(golden_dir).items():
        if len(d["start"]) != 2:  # simple_single / multiplayer_schelling: not 2-player, featurize_state cannot run
            continue
        n += 1
        L = oc.multienv_layout(golden_dir, name)
        assert L.GW * L.GH <= oc.OC_MAX_CELLS
        # only two layouts have tomato dispensers (featurize_state raises on a held tomato there)
        assert (sum(r.count("T") for r in d["grid"]) == 0) == (name not in ("mdp_test", "simple_tomato")), name
        feats, *_ = oc.replay(L, np.zeros((1, 2), np.uint8) + 4)
        assert feats.shape == (2, 2, 62) and np.isfinite(feats).all()
    assert n >= 16
