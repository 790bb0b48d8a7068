"""Recorded trajectories cut out of the vectorised rollout buffers (pantheonrl_b200/vec_record.py) against what
the reference's OWN recorder wrappers stored for the same scripted games (tests/golden/wrappers.npz: recorder
around LiarEnv; vec_record_rps.npz: recorder around RPSEnv).  The buffers come from the CPU oracle's scripted
rollout, whose layout and contents the device rollout equals bit for bit (GPU tests)."""
import io
import os

import numpy as np

import oracle
from oracle import rollout as orc
from pantheonrl_b200 import vec_record as vr
from pantheonrl_b200.common.trajsaver import SimultaneousTransitions, TurnBasedTransitions
from pantheonrl_b200.spaces import Discrete, MultiDiscrete

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _pad4(a):
    a = np.asarray(a).reshape(len(a), -1)
    out = np.zeros((a.shape[0], 4), np.uint8)
    out[:, :a.shape[1]] = a
    return out


def _npy(tr):
    b = io.BytesIO()
    tr.write_transition(b)
    return np.frombuffer(b.getvalue(), np.uint8)


def test_turn_based_recording_from_buffers_matches_the_reference_recorder():
    g = np.load(os.path.join(GOLD, "wrappers.npz"))
    T = 60
    ego_s, alt_s = g["liar_ego_script"], g["liar_alt_script"]
    ego_act = _pad4(np.array([ego_s[t % len(ego_s)] for t in range(T)]))
    alt_act = _pad4(np.array([alt_s[k % len(alt_s)] for k in range(3 * T)]))
    space = oracle.make_space(**oracle.LIAR_SPACE)
    ego, alt, carry = orc.rollout("liar", space, None, None, N=1, T=T, script_ego_act=ego_act, script_alt_act=alt_act,
                                  script_reset=g["liar_rf_resets"].astype(np.uint8), alt=orc.new_buffer(3 * T, 1, True))
    alt["count"] = alt["count"] + ((carry["flags"] >> 2) & 1)  # + the partner row left open at the end, if any
    tr = vr.turn_based_transitions(ego, alt, 0, carry["ego_last_done"][0])
    assert isinstance(tr, TurnBasedTransitions)
    assert np.array_equal(tr.obs, g["liar_rf_rec_obs"]) and np.array_equal(tr.acts, g["liar_rf_rec_acts"])
    assert np.array_equal(tr.flags, g["liar_rf_rec_flags"])
    assert np.array_equal(_npy(tr), g["liar_rf_npy"])  # the .npy file, byte for byte
    assert len(tr.get_ego_transitions()) == T and len(tr.get_alt_transitions()) == int(g["liar_rf_alt_n"])
    back = TurnBasedTransitions.read_transition(io.BytesIO(_npy(tr).tobytes()), MultiDiscrete([7] * 6 + [7, 12] * 12),
                                                MultiDiscrete([7, 12]))
    assert np.array_equal(back.flags, tr.flags)


def test_simultaneous_recording_from_buffers_matches_the_reference_recorder():
    g = np.load(os.path.join(GOLD, "vec_record_rps.npz"))
    T = int(g["T"])
    ego_act = _pad4(np.array([g["ego_script"][t % len(g["ego_script"])] for t in range(T)]))
    alt_act = _pad4(np.array([g["alt_script"][t % len(g["alt_script"])] for t in range(T)]))
    space = oracle.make_space(**oracle.RPS_SPACE)
    ego, alt, carry = orc.rollout("rps", space, None, None, N=1, T=T, script_ego_act=ego_act, script_alt_act=alt_act,
                                  alt=orc.new_buffer(T, 1, True))
    tr = vr.simultaneous_transitions(ego, alt, 0, carry["ego_last_done"][0], obs_len=1)
    assert isinstance(tr, SimultaneousTransitions)
    for name in ("egoobs", "egoacts", "altobs", "altacts", "flags"):
        assert np.array_equal(getattr(tr, name), g[name]), name
    assert np.array_equal(_npy(tr), g["npy"])
    back = SimultaneousTransitions.read_transition(io.BytesIO(_npy(tr).tobytes()), Discrete(1), Discrete(3))
    assert np.array_equal(back.altacts.reshape(-1), tr.altacts)
