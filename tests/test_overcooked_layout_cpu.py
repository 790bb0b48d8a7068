"""CPU test of host logic: pth_overcooked_layout_init (the product's table builder, runs on the
host, no device work) against the oracle's MotionPlanner restatement on every layout: for each
(floor cell, orientation) the closest onion / dish / serving deltas and the per-pot plan lengths
must give the same features the oracle derives from its all-pairs distance table."""
import ctypes as C

import numpy as np
import pytest

from oracle import overcooked as ooc
from pantheonrl_b200 import _lib
from pantheonrl_b200.envs import overcooked as oc


@pytest.mark.parametrize("layout", oc.LAYOUT_LIST)
def test_layout_tables_match_oracle(golden_dir, layout):
    L = oc.build_layout(layout)
    oL = ooc.multienv_layout(golden_dir, layout)
    W, Hh = L.width, L.height
    assert (W, Hh) == (oL.GW, oL.GH)
    sd = np.frombuffer(L.static_delta, np.int8).reshape(-1, 3, 2)
    pdist = np.frombuffer(L.pot_dist, np.uint8).reshape(-1, _lib.PTH_OC_MAX_POTS)
    lib = __import__("oracle").lib()
    f0 = np.zeros(62, np.float32)
    f1 = np.zeros(62, np.float32)
    floor = [(x, y) for y in range(Hh) for x in range(W) if L.terrain[y * W + x] == _lib.PTH_OC_FLOOR]
    other = floor[-1]
    for (x, y) in floor:
        for o in range(4):
            s = ooc.OrcOcState()
            ox, oy = other if other != (x, y) else floor[0]
            s.p[0:7] = (C.c_uint8 * 7)(x, y, o, 0, 0, 0, 0)
            s.p[7:14] = (C.c_uint8 * 7)(ox, oy, 0, 0, 0, 0, 0)
            s.order_len = 255
            assert lib.orc_oc_featurize(C.byref(oL), C.byref(s), f0.ctypes.data_as(C.c_void_p),
                                        f1.ctypes.data_as(C.c_void_p)) == 0
            node = (y * W + x) * 4 + o
            # player-0 block: 7,8 onion | 9,10 empty pot | 19,20 dish | 23,24 serving | 25..28 walls
            assert tuple(f0[7:9]) == tuple(sd[node, 0]), (layout, x, y, o, "onion")
            assert tuple(f0[19:21]) == tuple(sd[node, 1]), (layout, x, y, o, "dish")
            assert tuple(f0[23:25]) == tuple(sd[node, 2]), (layout, x, y, o, "serve")
            walls = [(L.wall[y * W + x] >> d) & 1 for d in range(4)]
            assert list(f0[25:29]) == walls
            # all pots are empty: the oracle's closest empty pot is the first pot with the smallest plan length
            d = pdist[node, :L.n_pots].astype(int)
            if (d < 255).any():
                p = int(np.argmin(d))
                assert tuple(f0[9:11]) == (L.pot_x[p] - x, L.pot_y[p] - y), (layout, x, y, o, "pot")
            else:
                assert tuple(f0[9:11]) == (0, 0)


def test_layout_rejects_bad_grids():
    with pytest.raises(ValueError):
        oc.build_layout(dict(grid=["XXTXX", "O 12O", "XDPSX"]))      # tomato dispenser
    with pytest.raises(_lib.PthError):
        oc.build_layout(dict(grid=["XXPXX", "O 12 ", "XDXSX"]))      # floor on the border
    with pytest.raises(ValueError):
        oc.build_layout(dict(grid=["XXPXX", "O 1 O", "XDXSX"]))      # one player only
