"""CPU ORACLE (test infrastructure): Python driver for orc_rollout (the C
restatement in pth_oracle_rollout.inc) — allocates numpy buffers shaped like the
device buffers of pth_rollout_args and returns them."""
import ctypes as C

import numpy as np

from . import OrcSpace, _p, lib


class OrcBuffer(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p),
                ("values", C.c_void_p), ("logp", C.c_void_p), ("episode_starts", C.c_void_p),
                ("count", C.c_void_p), ("Tcap", C.c_int64)]


class OrcRolloutArgs(C.Structure):
    _fields_ = [
        ("env_kind", C.c_int32), ("partner_records", C.c_int32),
        ("space", C.POINTER(OrcSpace)),
        ("ego_params", C.c_void_p), ("alt_params", C.c_void_p),
        ("ego", OrcBuffer), ("alt", OrcBuffer),
        ("ego_last_start", C.c_void_p), ("alt_last_done", C.c_void_p), ("total_rew", C.c_void_p),
        ("flags", C.c_void_p), ("game_state", C.c_void_p), ("ego_last_value", C.c_void_p),
        ("ego_last_done", C.c_void_p), ("ep_stats", C.c_void_p), ("alt_boot_done", C.c_void_p),
        ("N", C.c_int64), ("T", C.c_int64), ("env0", C.c_int64),
        ("seed", C.c_uint64), ("tick0", C.c_uint32), ("probegostart", C.c_float),
        ("first_rollout", C.c_int32),
        ("script_ego_act", C.c_void_p), ("script_alt_act", C.c_void_p), ("script_reset", C.c_void_p),
        ("oc_layout", C.c_void_p), ("oc_state", C.c_void_p), ("oc_ego_idx", C.c_int32),
    ]


OC_STATE_BYTES = 14 + 128 * 4 + 1 + 8 + 1 + 4  # sizeof(orc_oc_state): 540 (t is 4-byte aligned)


def alt_capacity(env_kind, T):
    """Rows of the partner's buffer: simultaneous games one per tick; Liar's Dice up to two per tick
    (reply + opening move after a reset) plus the open row carried in from the previous rollout."""
    return 2 * T + 1 if env_kind == "liar" else T


def new_buffer(Tcap, N, ragged, box=False):
    obs = np.zeros((Tcap, N, 64), np.float32) if box else np.zeros((Tcap, N, 32), np.uint8)
    b = dict(obs=obs, actions=np.zeros((Tcap, N, 4), np.uint8),
             rewards=np.zeros((Tcap, N), np.float32), values=np.zeros((Tcap, N), np.float32),
             logp=np.zeros((Tcap, N), np.float32), episode_starts=np.zeros((Tcap, N), np.float32))
    b["count"] = np.zeros(N, np.int32) if ragged else None
    return b


def new_carry(N):
    return dict(ego_last_start=np.ones(N, np.float32), alt_last_done=np.ones(N, np.float32),
                total_rew=np.zeros((2, N), np.float32), flags=np.zeros(N, np.uint8),
                game_state=np.zeros((N, 32), np.uint8), oc_state=np.zeros((N, OC_STATE_BYTES), np.uint8),
                ego_last_value=np.zeros(N, np.float32),
                ego_last_done=np.zeros(N, np.float32), ep_stats=np.zeros(4, np.float32),
                alt_boot_done=np.zeros(N, np.float32))


def _cbuf(b):
    o = OrcBuffer()
    for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts", "count"):
        setattr(o, k, None if b[k] is None else b[k].ctypes.data)
    o.Tcap = b["obs"].shape[0]
    return o


def rollout(env_kind, space, ego_params, alt_params, N, T, seed=10, tick0=0, env0=0,
            probegostart=0.5, first_rollout=True, partner_records=True, carry=None, ego=None,
            alt=None, script_ego_act=None, script_alt_act=None, script_reset=None, oc_layout=None,
            oc_ego_idx=0):
    """Runs one rollout; returns (ego buffer dict, partner buffer dict, carry dict)."""
    box = env_kind == "overcooked"
    ego = ego or new_buffer(T, N, False, box)
    alt = alt or new_buffer(alt_capacity(env_kind, T), N, True, box)
    carry = carry or new_carry(N)
    a = OrcRolloutArgs()
    a.env_kind = {"rps": 0, "liar": 1, "overcooked": 2}[env_kind]
    if box:
        assert oc_layout is not None
        a.oc_layout = C.addressof(oc_layout)
        a.oc_state = carry["oc_state"].ctypes.data
        a.oc_ego_idx = int(oc_ego_idx)
    a.partner_records = int(partner_records)
    a.space = C.pointer(space)
    keep = []
    for name, p in (("ego_params", ego_params), ("alt_params", alt_params)):
        if p is not None:
            p = np.ascontiguousarray(p, np.float32)
            keep.append(p)
            setattr(a, name, p.ctypes.data)
    a.ego, a.alt = _cbuf(ego), _cbuf(alt)
    for k in ("ego_last_start", "alt_last_done", "total_rew", "flags", "game_state",
              "ego_last_value", "ego_last_done", "ep_stats", "alt_boot_done"):
        setattr(a, k, carry[k].ctypes.data)
    a.N, a.T, a.env0 = N, T, env0
    a.seed, a.tick0, a.probegostart = seed, tick0, probegostart
    a.first_rollout = int(first_rollout)
    for name, s in (("script_ego_act", script_ego_act), ("script_alt_act", script_alt_act),
                    ("script_reset", script_reset)):
        if s is not None:
            s = np.ascontiguousarray(s, np.uint8)
            keep.append(s)
            setattr(a, name, s.ctypes.data)
    if alt["count"] is not None and first_rollout:
        alt["count"][:] = 0  # later rollouts read count[n]: where the carried open row sits
    lib().orc_rollout(C.byref(a))
    return ego, alt, carry


def partner_rows_from_events(g):
    """Rebuild what OnPolicyAgent would have stored (agents.py:172-179, 196-198)
    from the recorded get_action / update event log of a golden routing trace."""
    rows = []
    latch = True  # agents.py:97 _last_episode_starts = [True]
    for i in range(len(g["ev_kind"])):
        if g["ev_kind"][i] == 0:
            rows.append(dict(obs=g["ev_obs"][i], act=g["ev_act"][i], rew=0.0, start=float(latch),
                             pid=int(g["ev_pid"][i])))
        else:
            rows[-1]["rew"] += float(g["ev_rew"][i])
            latch = bool(g["ev_done"][i])
    return rows, latch
