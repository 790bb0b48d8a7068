"""Device rollout buffers + the binding of pth_rollout_run.

``RolloutBuffers`` owns the PyTorch tensors (HBM layout of DESIGN.md §2) that
the rollout megakernel fills; ``run_rollout`` launches one T-tick rollout.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import Context, check, current_stream

ENV_KINDS = {"rps": _lib.PTH_ENV_RPS, "liar": _lib.PTH_ENV_LIAR, "overcooked": _lib.PTH_ENV_OVERCOOKED}


def space_for(env_kind):
    if env_kind == "rps":
        return _lib.Space.onehot([1], [3])
    if env_kind == "liar":
        return _lib.Space.onehot([7] * 6 + [7, 12] * 12, [7, 12])
    if env_kind == "overcooked":
        return _lib.Space.box(_lib.PTH_OC_OBS, [6])
    raise ValueError(env_kind)


def alt_capacity(env_kind, T):
    """Rows of the partner's buffer: one per tick in simultaneous games; in Liar's Dice up to two per tick
    (reply + opening move after a reset) plus the open row carried in from the previous rollout."""
    return 2 * T + 1 if env_kind == "liar" else T


class Buffer:
    """One learner's rollout buffer: arrays [Tcap, N] (time-major, env contiguous)."""

    def __init__(self, Tcap, N, ragged, device, box=False):
        self.Tcap, self.N, self.ragged, self.box = Tcap, N, ragged, box
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=device)  # noqa: E731
        # observation rows: 32 bytes (one-hot slots) or 64 fp32 (Box)
        self.obs = z(Tcap, N, _lib.PTH_OC_ROW) if box else z(Tcap, N, 32, dt=torch.uint8)
        self.actions = z(Tcap, N, 4, dt=torch.uint8)
        self.rewards = z(Tcap, N)
        self.values = z(Tcap, N)
        self.logp = z(Tcap, N)
        self.episode_starts = z(Tcap, N)
        self.advantages = z(Tcap, N)
        self.returns = z(Tcap, N)
        self.count = z(N, dt=torch.int32) if ragged else None

    def c_struct(self):
        b = _lib.Buffer()
        b.d_obs = self.obs.data_ptr()
        b.d_actions = self.actions.data_ptr()
        b.d_rewards = self.rewards.data_ptr()
        b.d_values = self.values.data_ptr()
        b.d_logp = self.logp.data_ptr()
        b.d_episode_starts = self.episode_starts.data_ptr()
        b.d_count = self.count.data_ptr() if self.count is not None else None
        b.Tcap = self.Tcap
        return b


class Carry:
    """Per-env driver state carried across rollouts (MultiAgentEnv fields)."""

    def __init__(self, N, device, state_bytes=32):
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=device)  # noqa: E731
        self.ego_last_start = torch.ones(N, device=device)
        self.alt_last_done = torch.ones(N, device=device)
        self.total_rew = z(2, N)
        self.flags = z(N, dt=torch.uint8)
        self.game_state = z(N, state_bytes, dt=torch.uint8)
        self.ego_last_value = z(N)
        self.ego_last_done = z(N)
        self.ep_stats = z(4)
        self.alt_boot_done = z(N)

    def c_struct(self):
        c = _lib.EnvCarry()
        for k in ("ego_last_start", "alt_last_done", "total_rew", "flags", "game_state",
                  "ego_last_value", "ego_last_done", "ep_stats", "alt_boot_done"):
            setattr(c, "d_" + k, getattr(self, k).data_ptr())
        return c


def run_rollout(env_kind, space, ego_params, alt_params, ego, alt, carry, T, seed, tick0,
                env0=0, probegostart=0.5, first_rollout=False, partner_records=True, d_layout=None):
    """One T-tick rollout of ego.N on-device envs (pth_rollout_run)."""
    a = _lib.RolloutArgs()
    a.env_kind = ENV_KINDS[env_kind]
    a.partner_records = int(partner_records)
    a.space = C.pointer(space)
    a.d_ego_params = ego_params.data_ptr()
    a.d_alt_params = alt_params.data_ptr()
    a.ego = ego.c_struct()
    if alt is not None:
        a.alt = alt.c_struct()
    a.carry = carry.c_struct()
    a.N, a.T, a.env0 = ego.N, int(T), int(env0)
    a.seed, a.tick0 = int(seed), int(tick0) & 0xffffffff
    a.probegostart = float(probegostart)
    a.first_rollout = int(first_rollout)
    a.d_layout = d_layout.data_ptr() if d_layout is not None else None
    dev = ego_params.device
    ctx = Context.get(dev.index if dev.index is not None else torch.cuda.current_device())
    check(_lib.load().pth_rollout_run(ctx.handle, C.byref(a), current_stream()), "pth_rollout_run")
    _lib.count_launch()
