"""CPU ORACLE (test infrastructure): ctypes driver for the Overcooked restatement in
oracle/pth_oracle_overcooked.inc (layout tables, replay of joint actions, featurisation)."""
import ctypes as C
import json
import os

import numpy as np

from . import lib

OC_MAX_CELLS = 128
OC_OBS = 62
ROW = 64  # floats per observation row in the rollout buffers (62 used)


class OrcOcLayout(C.Structure):
    _fields_ = [
        ("GW", C.c_int32), ("GH", C.c_int32),
        ("cook_time", C.c_int32), ("num_items", C.c_int32), ("delivery_reward", C.c_int32),
        ("horizon", C.c_int32),
        ("rew_placement_in_pot", C.c_int32), ("rew_dish_pickup", C.c_int32), ("rew_soup_pickup", C.c_int32),
        ("start", (C.c_int32 * 2) * 2),
        ("n_order", C.c_int32), ("order", C.c_uint8 * 8),
        ("terrain", C.c_char * OC_MAX_CELLS),
        ("dist", (C.c_int16 * (OC_MAX_CELLS * 4)) * (OC_MAX_CELLS * 4)),
    ]


class OrcOcState(C.Structure):
    _fields_ = [("p", C.c_uint8 * 14), ("cell", C.c_uint8 * (OC_MAX_CELLS * 4)),
                ("order_len", C.c_uint8), ("order", C.c_uint8 * 8), ("t", C.c_int32)]


# OvercookedMultiEnv's constants (overcooked.py:18-28)
MULTIENV_SHAPING = dict(rew_placement_in_pot=3, rew_dish_pickup=3, rew_soup_pickup=5)


def make_layout(grid, start, cook_time=20, num_items=3, delivery_reward=20, horizon=400,
                order_list=None, rew_placement_in_pot=0, rew_dish_pickup=0, rew_soup_pickup=0):
    """grid: list of rows with player digits already removed (terrain_mtx)."""
    L = OrcOcLayout()
    L.GH, L.GW = len(grid), len(grid[0])
    assert L.GW * L.GH <= OC_MAX_CELLS and all(len(r) == L.GW for r in grid)
    L.cook_time, L.num_items, L.delivery_reward, L.horizon = cook_time, num_items, delivery_reward, horizon
    L.rew_placement_in_pot, L.rew_dish_pickup, L.rew_soup_pickup = \
        rew_placement_in_pot, rew_dish_pickup, rew_soup_pickup
    for i in range(2):
        L.start[i][0], L.start[i][1] = int(start[i][0]), int(start[i][1])
    if order_list is None:
        L.n_order = -1
    else:
        L.n_order = len(order_list)
        for i, o in enumerate(order_list):
            L.order[i] = int(o)
    L.terrain = "".join(grid).encode().ljust(OC_MAX_CELLS, b"\0")
    lib().orc_oc_layout_init(C.byref(L))
    return L


def named_layouts(golden_dir):
    return json.loads(str(np.load(os.path.join(golden_dir, "oc_layouts.npz"))["layouts"]))


def multienv_layout(golden_dir, name, horizon=400):
    """The layout as OvercookedMultiEnv(layout_name=name) configures it."""
    d = named_layouts(golden_dir)[name]
    return make_layout(d["grid"], d["start"], d["cook_time"], d["num_items"], d["delivery_reward"],
                       horizon, None, **MULTIENV_SHAPING)


def state_bytes(L):
    return 14 + 4 * L.GW * L.GH + 9


def replay(L, actions, want_states=False):
    """actions [S][2] in PLAYER order.  Returns feats [S+1][2][62], sparse [S], shaped [S], dones [S],
    states [S+1][state_bytes] (or None)."""
    actions = np.ascontiguousarray(actions, np.uint8)
    S = actions.shape[0]
    feats = np.zeros((S + 1, 2, OC_OBS), np.float32)
    sparse = np.zeros(S, np.int32)
    shaped = np.zeros(S, np.int32)
    dones = np.zeros(S, np.uint8)
    nb = state_bytes(L)
    states = np.zeros((S + 1, nb), np.uint8) if want_states else None
    lib().orc_oc_replay(C.byref(L), actions.ctypes.data_as(C.c_void_p), C.c_int64(S),
                        feats.ctypes.data_as(C.c_void_p), sparse.ctypes.data_as(C.c_void_p),
                        shaped.ctypes.data_as(C.c_void_p), dones.ctypes.data_as(C.c_void_p),
                        None if states is None else states.ctypes.data_as(C.c_void_p), C.c_int64(nb))
    return feats, sparse, shaped, dones, states
