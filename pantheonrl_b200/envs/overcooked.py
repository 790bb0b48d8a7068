"""Overcooked behind the reference's OvercookedMultiEnv API
(overcookedgym/overcooked.py:10-98), with the gridworld rules, the planner-distance
features and the reward shaping running in CUDA (csrc/pth_overcooked.cuh).

``OvercookedMultiEnv(layout_name, ego_agent_idx=0)`` is a SimultaneousEnv: N = 1 calls
(``multi_step`` / ``multi_reset``) hit pth_env_overcooked_step / _reset; with n_envs > 1 the
whole collect-rollouts loop runs inside pth_rollout_run (engine.VecTrainer("overcooked", ...)).

Layout grids: the two-player onion layouts of the reference's LAYOUT_LIST
(overcooked_ai_py/data/layouts/*.layout), '1' / '2' marking the start cells.  The two tomato
layouts (mdp_test, simple_tomato) are not offered: the reference's own featurize_state raises on
a held tomato (overcooked_mdp.py:1096).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._lib import Context, check, current_stream
from ..common.multiagentenv import SimultaneousEnv
from ..spaces import Box, Discrete

LAYOUTS = {
    'corridor': dict(grid=['XXXXXOXXDXXXXX', 'X  1  XX  2  X', 'X     XX     X', 'X  XXXXXXXX  X', 'X            X', 'X  XXXXXXXX  X', 'X     XX     X', 'X     XX     X', 'XXXXXSXXPPXXXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'five_by_five': dict(grid=['XDPXX', 'X   S', 'O 2 X', 'X1  D', 'XOXPX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'random0': dict(grid=['XXXPX', 'O X1P', 'O2X X', 'D X X', 'XXXSX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'random1': dict(grid=['XXXPX', 'X 1 P', 'D2X X', 'O   X', 'XOSXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'random2': dict(grid=['XXXPX', 'O X1P', 'O2X X', 'D X X', 'XXXSX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'random3': dict(grid=['XXXPPXXX', 'X  2   X', 'D XXXX S', 'X  1   X', 'XXXOOXXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'scenario1_s': dict(grid=['XXOXDXX', 'X 1X2 X', 'X  X  X', 'X     X', 'XSXXPPX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'scenario2': dict(grid=['XXXXXOXXXX', 'S        O', 'D    1 2 X', 'XXXXXXPXXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'scenario2_s': dict(grid=['XXOXXXX', 'S     O', 'D 1 2 X', 'XXXPXXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'scenario3': dict(grid=['XXXXXOXXXX', 'S     XXPX', 'X    1   X', 'D XXXXXX X', 'X     2  O', 'XXXXXXXXXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'scenario4': dict(grid=['XXXXXOXXXX', 'S      XPX', 'D    1   X', 'XXXXXXXX X', 'XXXXXX2  O', 'XXXXXXXXXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'schelling': dict(grid=['XXSPDXX', 'X  1  X', 'X  X  X', 'O     O', 'X  X  X', 'X  2  X', 'XXDPSXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'schelling_s': dict(grid=['XSPDX', 'X 1 X', 'O   O', 'X 2 X', 'XDPSX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'simple': dict(grid=['XXPXX', 'O  2O', 'X1  X', 'XDXSX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'small_corridor': dict(grid=['XXXXXOXDXXXXX', 'X  1  X  2  X', 'X  XXXXXXX  X', 'X           X', 'XSXXXXXXXXPPX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'unident': dict(grid=['XXXXXXXXXXX', 'O XXSXOXX S', 'X    P  1 X', 'X2   P    X', 'XXXXDXDXXXX'],
        cook_time=20, num_items=3, delivery_reward=20),
    'unident_s': dict(grid=['XXXXXXXXX', 'O XSXOX S', 'X   P 1 X', 'X2  P   X', 'XXXDXDXXX'],
        cook_time=20, num_items=3, delivery_reward=20),
}

LAYOUT_LIST = sorted(LAYOUTS)
# overcookedgym/overcooked_utils.py:7-13: the names the Overcooked-AI papers use for five layouts
NAME_TRANSLATION = {"cramped_room": "simple", "asymmetric_advantages": "unident_s", "coordination_ring": "random1",
                    "forced_coordination": "random0", "counter_circuit": "random3"}
TERRAIN_CODE = {" ": _lib.PTH_OC_FLOOR, "X": _lib.PTH_OC_COUNTER, "O": _lib.PTH_OC_ONION, "P": _lib.PTH_OC_POT,
                "D": _lib.PTH_OC_DISH, "S": _lib.PTH_OC_SERVE}
# OvercookedMultiEnv's constants (overcooked.py:18-28)
HORIZON = 400
REW_SHAPING = dict(rew_placement_in_pot=3, rew_dish_pickup=3, rew_soup_pickup=5)
ACTION_NAMES = ["NORTH", "SOUTH", "EAST", "WEST", "STAY", "INTERACT"]  # Action.INDEX_TO_ACTION


def build_layout(layout, ego_agent_idx=0, horizon=HORIZON, **overrides):
    """A filled-in pth_overcooked_layout (host struct) for a layout name or a dict with
    ``grid`` rows ('1'/'2' = start cells), ``cook_time``, ``num_items``, ``delivery_reward``."""
    spec = dict(LAYOUTS[layout]) if isinstance(layout, str) else dict(layout)
    spec.update(overrides)
    grid = [str(r) for r in spec["grid"]]
    L = _lib.OvercookedLayout()
    L.height, L.width = len(grid), len(grid[0])
    if any(len(r) != L.width for r in grid) or L.width * L.height > _lib.PTH_OC_MAX_CELLS:
        raise ValueError("ragged grid or more than 128 cells")
    L.cook_time, L.num_items = int(spec.get("cook_time", 20)), int(spec.get("num_items", 3))
    L.delivery_reward, L.horizon = int(spec.get("delivery_reward", 20)), int(horizon)
    for k, v in REW_SHAPING.items():
        setattr(L, k, int(spec.get(k, v)))
    L.ego_agent_idx = int(ego_agent_idx)
    starts = {}
    for y, row in enumerate(grid):
        for x, ch in enumerate(row):
            if ch in "12":
                starts[int(ch) - 1] = (x, y)
                ch = " "
            if ch not in TERRAIN_CODE:
                raise ValueError(f"terrain {ch!r} is not supported (onion layouts only)")
            L.terrain[y * L.width + x] = TERRAIN_CODE[ch]
    if "start" in spec:
        starts = {i: tuple(p) for i, p in enumerate(spec["start"])}
    if sorted(starts) != [0, 1]:
        raise ValueError("the grid needs exactly the start cells '1' and '2'")
    for i in range(2):
        L.start_x[i], L.start_y[i] = starts[i]
    check(_lib.load().pth_overcooked_layout_init(C.byref(L)), "pth_overcooked_layout_init")
    return L


def layout_to_device(L, device="cuda"):
    """Device copy of an initialised layout (the pointer pth_rollout_run / pth_env_overcooked_* take)."""
    raw = np.frombuffer(bytes(L), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)


def space():
    """Box(62) observations, Discrete(6) actions (overcooked.py:38-49)."""
    return _lib.Space.box(_lib.PTH_OC_OBS, [6])


def env_reset(d_layout, N, device="cuda"):
    state = torch.zeros(N, _lib.PTH_OC_STATE_BYTES, dtype=torch.uint8, device=device)
    obs = torch.empty(N, 2, _lib.PTH_OC_ROW, dtype=torch.float32, device=device)
    ctx = Context.get(torch.device(device).index or 0)
    check(_lib.load().pth_env_overcooked_reset(ctx.handle, d_layout.data_ptr(), state.data_ptr(), obs.data_ptr(),
                                               N, current_stream()), "pth_env_overcooked_reset")
    _lib.count_launch()
    return state, obs


def env_step(d_layout, state, ego_action, alt_action):
    """state [N, 40] u8 updated in place; actions u8 [N]. Returns obs [N,2,64], reward [N], done [N]."""
    N = state.shape[0]
    dev = state.device
    obs = torch.empty(N, 2, _lib.PTH_OC_ROW, dtype=torch.float32, device=dev)
    rew = torch.empty(N, dtype=torch.float32, device=dev)
    done = torch.empty(N, dtype=torch.uint8, device=dev)
    ctx = Context.get(dev.index if dev.index is not None else torch.cuda.current_device())
    check(_lib.load().pth_env_overcooked_step(ctx.handle, d_layout.data_ptr(), state.data_ptr(),
                                              ego_action.data_ptr(), alt_action.data_ptr(), obs.data_ptr(),
                                              rew.data_ptr(), done.data_ptr(), N, current_stream()),
          "pth_env_overcooked_step")
    _lib.count_launch()
    return obs, rew, done


class OvercookedMultiEnv(SimultaneousEnv):
    """Drop-in for overcookedgym.overcooked.OvercookedMultiEnv (layout_name, ego_agent_idx)."""

    device_kind = "overcooked"

    def __init__(self, layout_name, ego_agent_idx=0, baselines=False, device="cuda"):
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.PthError("OvercookedMultiEnv steps on the GPU: no CUDA device")
        self.layout_name, self.ego_agent_idx, self.device = layout_name, int(ego_agent_idx), device
        self.layout = build_layout(layout_name, ego_agent_idx)
        self.d_layout = layout_to_device(self.layout, device)
        high = np.full((_lib.PTH_OC_OBS,), np.inf, dtype=np.float32)
        self.observation_space = Box(-high, high, dtype=np.float64)
        self.lA = 6
        self.action_space = Discrete(self.lA)
        self.multi_reset()

    def _obs_pair(self, obs):
        o = obs[0, :, :_lib.PTH_OC_OBS].cpu().numpy().astype(np.float64)
        return o[0], o[1]

    def multi_step(self, ego_action, alt_action):
        ea = torch.tensor([int(ego_action)], dtype=torch.uint8, device=self.device)
        aa = torch.tensor([int(alt_action)], dtype=torch.uint8, device=self.device)
        obs, rew, done = env_step(self.d_layout, self.state, ea, aa)
        r = float(rew.item())
        return self._obs_pair(obs), (r, r), bool(done.item()), {}

    def multi_reset(self):
        self.state, obs = env_reset(self.d_layout, 1, self.device)
        return self._obs_pair(obs)

    def render(self, mode="human", close=False):
        pass
