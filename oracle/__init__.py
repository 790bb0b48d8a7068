"""CPU ORACLE bindings (test infrastructure, not product code).

ctypes wrapper over oracle/_ref/libpth_oracle.so (built from pth_oracle.c by
oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libpth_oracle.so")
_lib = None


class OrcSpace(C.Structure):
    _fields_ = [
        ("obs_kind", C.c_int32),
        ("obs_len", C.c_int32),
        ("obs_nvec", C.c_int32 * 96),
        ("n_heads", C.c_int32),
        ("head_n", C.c_int32 * 4),
    ]


def make_space(nvec=None, heads=(3,), box_dim=None):
    s = OrcSpace()
    if box_dim is not None:
        s.obs_kind, s.obs_len = 1, int(box_dim)
    else:
        s.obs_kind, s.obs_len = 0, len(nvec)
        for i, v in enumerate(nvec):
            s.obs_nvec[i] = int(v)
    s.n_heads = len(heads)
    for i, v in enumerate(heads):
        s.head_n[i] = int(v)
    return s


RPS_SPACE = dict(nvec=[1], heads=[3])
LIAR_NVEC = [7] * 6 + [7, 12] * 12
LIAR_SPACE = dict(nvec=LIAR_NVEC, heads=[7, 12])
LIAR3_SPACE = dict(nvec=LIAR_NVEC * 3, heads=[7, 12])  # frame_wrap(LiarEnv(), 3): 90 slots, rows of 96 bytes


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
        _lib.orc_param_count.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def philox_raw(ctr, key):
    out = np.zeros(4, np.uint32)
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    lib().orc_philox_raw(_p(c), _p(k), _p(out))
    return out


def philox(seed, stream, index, tick, slot):
    out = np.zeros(4, np.uint32)
    lib().orc_philox(C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(index), C.c_uint32(tick),
                     C.c_uint32(slot), _p(out))
    return out


def math_vec(which, x):
    x = _f32(x)
    y = np.empty_like(x)
    lib().orc_math_vec(C.c_int({"exp": 0, "log": 1, "tanh": 2}[which]), _p(x), _p(y),
                       C.c_int64(x.size))
    return y


def gae(rew, val, start, last_values, dones, gamma=0.99, lam=0.95):
    rew, val, start = _f32(rew), _f32(val), _f32(start)
    last_values, dones = _f32(last_values), _f32(dones)
    T, N = rew.shape
    adv = np.empty((T, N), np.float32)
    ret = np.empty((T, N), np.float32)
    lib().orc_gae(_p(rew), _p(val), _p(start), _p(last_values), _p(dones), _p(adv), _p(ret),
                  C.c_int64(T), C.c_int64(N), C.c_double(gamma), C.c_double(lam))
    return adv, ret


def gae_ragged(rew, val, start, count, last_done, gamma=0.99, lam=0.95):
    rew, val, start = _f32(rew), _f32(val), _f32(start)
    count = np.ascontiguousarray(count, np.int32)
    last_done = _f32(last_done)
    T, N = rew.shape
    adv = np.zeros((T, N), np.float32)
    ret = np.zeros((T, N), np.float32)
    lib().orc_gae_ragged(_p(rew), _p(val), _p(start), _p(count), _p(last_done), _p(adv), _p(ret),
                         C.c_int64(T), C.c_int64(N), C.c_double(gamma), C.c_double(lam))
    return adv, ret


def rps_step(ego_a, alt_a):
    ego_a = np.ascontiguousarray(ego_a, np.int32)
    alt_a = np.ascontiguousarray(alt_a, np.int32)
    re = np.empty(ego_a.shape, np.float32)
    ra = np.empty(ego_a.shape, np.float32)
    lib().orc_rps_step(_p(ego_a), _p(alt_a), _p(re), _p(ra), C.c_int64(ego_a.size))
    return re, ra


def liar_reset(N, seed, tick, env0=0, probegostart=0.5):
    state = np.zeros((N, 32), np.uint8)
    ego_first = np.zeros(N, np.uint8)
    obs = np.zeros((N, 32), np.uint8)
    lib().orc_liar_reset(_p(state), _p(ego_first), _p(obs), C.c_int64(N), C.c_uint64(seed),
                         C.c_uint32(tick), C.c_int64(env0), C.c_float(probegostart))
    return state, ego_first, obs


def liar_step(state, is_ego, action):
    """state is modified in place. Returns obs, r_ego, r_alt, done."""
    assert state.dtype == np.uint8 and state.flags.c_contiguous
    N = state.shape[0]
    is_ego = np.ascontiguousarray(is_ego, np.uint8)
    action = np.ascontiguousarray(action, np.uint8)
    obs = np.zeros((N, 32), np.uint8)
    re = np.empty(N, np.float32)
    ra = np.empty(N, np.float32)
    done = np.zeros(N, np.uint8)
    lib().orc_liar_step(_p(state), _p(is_ego), _p(action), _p(obs), _p(re), _p(ra), _p(done),
                        C.c_int64(N))
    return obs, re, ra, done


def param_count(space):
    return int(lib().orc_param_count(C.byref(space)))


def policy_forward(space, params, obs, seed=0, rng_stream=2, tick=0, slot=0, idx0=0,
                   action_in=None, want_logits=True):
    params = _f32(params)
    if space.obs_kind == 0:
        obs = np.ascontiguousarray(obs, np.uint8)
    else:
        obs = _f32(obs)
    B, stride = obs.shape
    L = sum(space.head_n[i] for i in range(space.n_heads))
    action = np.zeros((B, 4), np.uint8)
    value = np.empty(B, np.float32)
    logp = np.empty(B, np.float32)
    ent = np.empty(B, np.float32)
    logits = np.empty((B, L), np.float32) if want_logits else None
    if action_in is not None:
        action_in = np.ascontiguousarray(action_in, np.uint8)
    lib().orc_policy_forward(C.byref(space), _p(params), _p(obs), C.c_int64(stride), C.c_int64(B),
                             C.c_uint64(seed), C.c_uint32(rng_stream), C.c_uint32(tick),
                             C.c_uint32(slot), C.c_int64(idx0), _p(action_in), _p(action),
                             _p(value), _p(logp), _p(ent), _p(logits))
    return dict(action=action, value=value, logp=logp, entropy=ent, logits=logits)


def adap_param_count(space, context_size):
    F = sum(space.obs_nvec[i] for i in range(space.obs_len)) if space.obs_kind == 0 else space.obs_len
    L = sum(space.head_n[i] for i in range(space.n_heads))
    return 2 * (64 * (F + context_size) + 64 + 64 * 64 + 64) + L * 64 + L + 64 + 1


def adap_forward(space, params, obs, ctx, seed=0, rng_stream=2, tick=0, slot=0, idx0=0, action_in=None):
    """AdapPolicy forward / evaluate_actions: ctx is [B, C] (one context per sample) or [C] (broadcast)."""
    params = _f32(params)
    obs = np.ascontiguousarray(obs, np.uint8) if space.obs_kind == 0 else _f32(obs)
    ctx = _f32(ctx)
    B, stride = obs.shape
    Cn = ctx.shape[-1]
    cstride = Cn if ctx.ndim == 2 else 0
    L = sum(space.head_n[i] for i in range(space.n_heads))
    action = np.zeros((B, 4), np.uint8)
    value, logp, ent = np.empty(B, np.float32), np.empty(B, np.float32), np.empty(B, np.float32)
    logits = np.empty((B, L), np.float32)
    if action_in is not None:
        action_in = np.ascontiguousarray(action_in, np.uint8)
    lib().orc_adap_forward(C.byref(space), _p(params), C.c_int32(Cn), _p(obs), C.c_int64(stride), _p(ctx),
                           C.c_int64(cstride), C.c_int64(B), C.c_uint64(seed), C.c_uint32(rng_stream),
                           C.c_uint32(tick), C.c_uint32(slot), C.c_int64(idx0), _p(action_in), _p(action),
                           _p(value), _p(logp), _p(ent), _p(logits))
    return dict(action=action, value=value, logp=logp, entropy=ent, logits=logits)


def adap_mult_param_count(space, context_size):
    lib().orc_adap_mult_param_count.restype = C.c_int64
    return int(lib().orc_adap_mult_param_count(C.byref(space), C.c_int32(context_size)))


def adap_mult_forward(space, params, obs, ctx, seed=0, rng_stream=2, tick=0, slot=0, idx0=0, action_in=None):
    """AdapPolicyMult forward / evaluate_actions: ctx is [B, C] (one context per sample) or [C] (broadcast)."""
    params = _f32(params)
    obs = np.ascontiguousarray(obs, np.uint8) if space.obs_kind == 0 else _f32(obs)
    ctx = _f32(ctx)
    B, stride = obs.shape
    Cn = ctx.shape[-1]
    cstride = Cn if ctx.ndim == 2 else 0
    L = sum(space.head_n[i] for i in range(space.n_heads))
    action = np.zeros((B, 4), np.uint8)
    value, logp, ent = np.empty(B, np.float32), np.empty(B, np.float32), np.empty(B, np.float32)
    logits = np.empty((B, L), np.float32)
    if action_in is not None:
        action_in = np.ascontiguousarray(action_in, np.uint8)
    lib().orc_adap_mult_forward(C.byref(space), _p(params), C.c_int32(Cn), _p(obs), C.c_int64(stride), _p(ctx),
                                C.c_int64(cstride), C.c_int64(B), C.c_uint64(seed), C.c_uint32(rng_stream),
                                C.c_uint32(tick), C.c_uint32(slot), C.c_int64(idx0), _p(action_in), _p(action),
                                _p(value), _p(logp), _p(ent), _p(logits))
    return dict(action=action, value=value, logp=logp, entropy=ent, logits=logits)
