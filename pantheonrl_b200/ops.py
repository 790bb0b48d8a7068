"""Tensor-level bindings: PyTorch tensors in, raw device pointers out.

Each function validates its tensors (CUDA, dtype, contiguous), allocates the
outputs with torch and calls exactly one C-ABI entry point of
libpantheon_b200.so on the current CUDA stream.  No math happens in Python.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import Context, check, current_stream, ptr


def _need(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (this path has no CPU implementation)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ctx(t):
    return Context.get(t.device.index if t.device.index is not None else torch.cuda.current_device())


# ----------------------------------------------------------------------- GAE
def gae(rewards, values, episode_starts, last_values, dones, gamma=0.99, gae_lambda=0.95,
        variant=0, out=None):
    """RolloutBuffer.compute_returns_and_advantage over [T, N] fp32 buffers."""
    f32 = torch.float32
    for n, t in (("rewards", rewards), ("values", values), ("episode_starts", episode_starts),
                 ("last_values", last_values), ("dones", dones)):
        _need(t, f32, n)
    T, N = rewards.shape
    if values.shape != (T, N) or episode_starts.shape != (T, N):
        raise ValueError("values / episode_starts must be [T, N]")
    if last_values.numel() != N or dones.numel() != N:
        raise ValueError("last_values / dones must have N elements")
    if out is None:
        adv = torch.empty_like(rewards)
        ret = torch.empty_like(rewards)
    else:
        adv, ret = out
    ctx = _ctx(rewards)
    check(_lib.load().pth_gae_f32(ctx.handle, ptr(rewards), ptr(values), ptr(episode_starts),
                                  ptr(last_values), ptr(dones), ptr(adv), ptr(ret), T, N,
                                  float(gamma), float(gae_lambda), int(variant), current_stream()),
          "pth_gae_f32")
    _lib.count_launch()
    return adv, ret


def gae_ragged(rewards, values, episode_starts, count, last_done, gamma=0.99, gae_lambda=0.95,
               out=None):
    f32 = torch.float32
    for n, t in (("rewards", rewards), ("values", values), ("episode_starts", episode_starts),
                 ("last_done", last_done)):
        _need(t, f32, n)
    _need(count, torch.int32, "count")
    T, N = rewards.shape
    if out is None:
        adv = torch.zeros_like(rewards)
        ret = torch.zeros_like(rewards)
    else:
        adv, ret = out
    ctx = _ctx(rewards)
    check(_lib.load().pth_gae_ragged_f32(ctx.handle, ptr(rewards), ptr(values), ptr(episode_starts),
                                         ptr(count), ptr(last_done), ptr(adv), ptr(ret), T, N,
                                         float(gamma), float(gae_lambda), current_stream()),
          "pth_gae_ragged_f32")
    _lib.count_launch()
    return adv, ret


# ----------------------------------------------------------------------- envs
def rps_step(ego_action, alt_action):
    _need(ego_action, torch.int32, "ego_action")
    _need(alt_action, torch.int32, "alt_action")
    N = ego_action.numel()
    re = torch.empty(N, dtype=torch.float32, device=ego_action.device)
    ra = torch.empty_like(re)
    check(_lib.load().pth_env_rps_step(_ctx(ego_action).handle, ptr(ego_action), ptr(alt_action),
                                       ptr(re), ptr(ra), N, current_stream()), "pth_env_rps_step")
    return re, ra


def liar_reset(N, seed, tick, env0=0, probegostart=0.5, device="cuda"):
    state = torch.zeros(N, 32, dtype=torch.uint8, device=device)
    ego_first = torch.zeros(N, dtype=torch.uint8, device=device)
    obs = torch.zeros(N, 32, dtype=torch.uint8, device=device)
    check(_lib.load().pth_env_liar_reset(_ctx(state).handle, ptr(state), ptr(ego_first), ptr(obs), N,
                                         int(seed), int(tick), int(env0), float(probegostart),
                                         current_stream()), "pth_env_liar_reset")
    return state, ego_first, obs


def liar_step(state, is_ego, action):
    """state [N, 32] u8 is updated in place; returns obs, r_ego, r_alt, done."""
    _need(state, torch.uint8, "state")
    _need(is_ego, torch.uint8, "is_ego")
    _need(action, torch.uint8, "action")
    N = state.shape[0]
    obs = torch.empty(N, 32, dtype=torch.uint8, device=state.device)
    re = torch.empty(N, dtype=torch.float32, device=state.device)
    ra = torch.empty_like(re)
    done = torch.empty(N, dtype=torch.uint8, device=state.device)
    check(_lib.load().pth_env_liar_step(_ctx(state).handle, ptr(state), ptr(is_ego), ptr(action),
                                        ptr(obs), ptr(re), ptr(ra), ptr(done), N, current_stream()),
          "pth_env_liar_step")
    return obs, re, ra, done


# ----------------------------------------------------------------------- policy
def policy_forward(space, params, obs, seed=0, rng_stream=_lib.STREAM_EGO, tick=0, slot=0, idx0=0,
                   action_in=None, want=("action", "value", "logp", "entropy", "logits"), race=None, context=None,
                   num_partners=0, partner_idx=0, out=None, adap_mult=False):
    """ActorCriticPolicy.forward (sampling) or evaluate_actions (action_in given).

    obs: [B, stride] uint8 (one-hot spaces) or float32 (Box). Returns a dict of
    tensors for the names in ``want``.  context: AdapPolicy's context inputs, float32 [B, C] (one per
    sample) or [C] / [1, C] (one for the whole batch); params then have the AdapPolicy layout."""
    _need(params, torch.float32, "params")
    if space.obs_kind == _lib.PTH_OBS_ONEHOT:
        _need(obs, torch.uint8, "obs")
    else:
        _need(obs, torch.float32, "obs")
    B, stride = obs.shape
    dev = obs.device
    L = sum(space.heads)
    out = dict(out) if out is not None else {}  # caller-owned output tensors (the N = 1 facade reuses one packed buffer)
    if "action" in want and "action" not in out:
        out["action"] = torch.zeros(B, 4, dtype=torch.uint8, device=dev)
    for k in ("value", "logp", "entropy"):
        if k in want and k not in out:
            out[k] = torch.empty(B, dtype=torch.float32, device=dev)
    if "logits" in want:
        out["logits"] = torch.empty(B, L, dtype=torch.float32, device=dev)
    if action_in is not None:
        _need(action_in, torch.uint8, "action_in")
    a = _lib.ForwardArgs()
    a.space = C.pointer(space)
    a.d_params = params.data_ptr()
    a.d_obs = obs.data_ptr()
    a.obs_stride = stride
    a.B = B
    a.seed = int(seed)
    a.rng_stream = int(rng_stream)
    a.tick = int(tick)
    a.slot = int(slot)
    a.idx0 = int(idx0)
    a.d_action_in = action_in.data_ptr() if action_in is not None else None
    a.d_action = out["action"].data_ptr() if "action" in out else None
    a.d_value = out["value"].data_ptr() if "value" in out else None
    a.d_logp = out["logp"].data_ptr() if "logp" in out else None
    a.d_entropy = out["entropy"].data_ptr() if "entropy" in out else None
    a.d_logits = out["logits"].data_ptr() if "logits" in out else None
    if race is not None:  # exponential(1) draws [B, L] from the host's generator (reference-RNG compatibility)
        _need(race, torch.float32, "race")
        if tuple(race.shape) != (B, L):
            raise ValueError("race must be [B, L]")
        a.d_race = race.data_ptr()
    a.num_partners, a.partner_idx = int(num_partners), int(partner_idx)  # ModularPolicy: partner module composing the outputs
    a.adap_mult = int(bool(adap_mult))  # AdapPolicyMult parameter layout (with context=)
    if context is not None:
        _need(context, torch.float32, "context")
        a.context_size, a.d_context = context.shape[-1], context.data_ptr()
        a.context_stride = context.shape[-1] if context.dim() == 2 and context.shape[0] == B and B > 1 else 0
        if context.dim() == 2 and context.shape[0] not in (1, B):
            raise ValueError("context must be [B, C], [1, C] or [C]")
    check(_lib.load().pth_policy_forward(_ctx(obs).handle, C.byref(a), current_stream()),
          "pth_policy_forward")
    _lib.count_launch()
    return out
