"""CPU ORACLE (test infrastructure): Python driver for orc_ppo_update,
orc_perm_feistel and orc_index_build (pth_oracle_update.inc)."""
import ctypes as C
import ctypes as C_

import numpy as np

from . import OrcSpace, lib


class OrcUpdateArgs(C.Structure):
    _fields_ = [
        ("space", C.POINTER(OrcSpace)),
        ("params", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
        ("adam_step", C.c_int64),
        ("obs", C.c_void_p), ("actions", C.c_void_p), ("old_logp", C.c_void_p),
        ("advantages", C.c_void_p), ("returns", C.c_void_p), ("index", C.c_void_p),
        ("perm", C.c_void_p),
        ("M", C.c_int64), ("batch_size", C.c_int64), ("n_epochs", C.c_int32),
        ("learning_rate", C.c_float), ("clip_range", C.c_float), ("ent_coef", C.c_float),
        ("vf_coef", C.c_float), ("max_grad_norm", C.c_float),
        ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("adam_eps", C.c_float),
        ("normalize_advantage", C.c_int32), ("grid", C.c_int32), ("world", C.c_int32),
        ("stats", C.c_void_p), ("grad_out", C.c_void_p),
        ("loss_kind", C.c_int32), ("l2_weight", C.c_float),
        ("context_size", C.c_int32), ("ctx", C.c_void_p), ("ctx_loss_coeff", C.c_float),
        ("n_ctx", C.c_int32), ("n_states", C.c_int32), ("ctx_sidx", C.c_void_p), ("ctx_draws", C.c_void_p), ("ctx_loss_out", C.c_void_p), ("adap_mult", C.c_int32),
    ]


def perm_feistel(M, n_epochs, seed, stream, epoch0=0):
    perm = np.zeros((n_epochs, M), np.int32)
    lib().orc_perm_feistel(perm.ctypes.data_as(C.c_void_p), C.c_int64(M), C.c_int32(n_epochs),
                           C.c_uint64(seed), C.c_uint32(stream), C.c_uint32(epoch0))
    return perm


def index_build(count, T, N):
    index = np.zeros(T * N, np.int32)
    total = np.zeros(1, np.int32)
    cp = None if count is None else np.ascontiguousarray(count, np.int32)
    lib().orc_index_build(None if cp is None else cp.ctypes.data_as(C.c_void_p), C.c_int64(T),
                          C.c_int64(N), index.ctypes.data_as(C.c_void_p),
                          total.ctypes.data_as(C.c_void_p))
    return index[:int(total[0])].copy()


def ppo_update(space, params, adam_m, adam_v, adam_step, obs, actions, old_logp, advantages,
               returns, perm, batch_size, grid, index=None, learning_rate=3e-4, clip_range=0.2,
               ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, betas=(0.9, 0.999), eps=1e-5,
               normalize_advantage=True, world=1, loss_kind=0, l2_weight=0.0, ctx=None, ctx_loss_coeff=0.0,
               ctx_sidx=None, ctx_draws=None, adap_mult=False):
    """Runs PPO.train on flat sample arrays; params / adam state are updated IN PLACE
    (float32 numpy arrays). Returns (stats [n_epochs*n_mb, 8], last pre-clip gradient)."""
    f32 = lambda x: np.ascontiguousarray(x, np.float32)  # noqa: E731
    for a in (params, adam_m, adam_v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    if space.obs_kind == 1:  # Box rows of 64 floats
        obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 64)
    else:
        obs = np.ascontiguousarray(obs, np.uint8).reshape(-1, 32 if space.obs_len <= 32 else 96)
    actions = np.ascontiguousarray(actions, np.uint8).reshape(-1, 4)
    old_logp, advantages, returns = f32(old_logp).reshape(-1), f32(advantages).reshape(-1), f32(returns).reshape(-1)
    perm = np.ascontiguousarray(perm, np.int32)
    n_epochs, M = perm.shape
    n_mb = (M + batch_size - 1) // batch_size
    stats = np.zeros((n_epochs * n_mb, 8), np.float32)
    grad = np.zeros(params.size, np.float32)
    a = OrcUpdateArgs()
    a.space = C.pointer(space)
    a.params, a.adam_m, a.adam_v = params.ctypes.data, adam_m.ctypes.data, adam_v.ctypes.data
    a.adam_step = int(adam_step)
    a.obs, a.actions = obs.ctypes.data, actions.ctypes.data
    a.old_logp, a.advantages, a.returns = old_logp.ctypes.data, advantages.ctypes.data, returns.ctypes.data
    if index is not None:
        index = np.ascontiguousarray(index, np.int32)
        a.index = index.ctypes.data
    a.perm = perm.ctypes.data
    a.M, a.batch_size, a.n_epochs = M, int(batch_size), n_epochs
    a.learning_rate, a.clip_range, a.ent_coef = learning_rate, clip_range, ent_coef
    a.vf_coef, a.max_grad_norm = vf_coef, max_grad_norm
    a.adam_beta1, a.adam_beta2, a.adam_eps = betas[0], betas[1], eps
    a.normalize_advantage = int(normalize_advantage)
    a.loss_kind, a.l2_weight = int(loss_kind), float(l2_weight)
    if ctx is not None:  # AdapPolicy: [rows][C] contexts stored with the samples
        ctx = f32(ctx)
        a.context_size, a.ctx = ctx.shape[1], ctx.ctypes.data
    if ctx_sidx is not None:  # ADAP context loss: [n_epochs * n_mb][S] positions, [n_epochs * n_mb][K][C] contexts
        ctx_sidx = np.ascontiguousarray(ctx_sidx, np.int32)
        ctx_draws = f32(ctx_draws)
        assert ctx_sidx.shape[0] == n_epochs * n_mb and ctx_draws.shape[0] == n_epochs * n_mb
        a.ctx_loss_coeff, a.n_states, a.n_ctx = float(ctx_loss_coeff), ctx_sidx.shape[1], ctx_draws.shape[1]
        a.ctx_sidx, a.ctx_draws = ctx_sidx.ctypes.data, ctx_draws.ctypes.data
        ctx_loss = np.zeros(n_epochs * n_mb, np.float32)
        a.ctx_loss_out = ctx_loss.ctypes.data
    a.adap_mult = int(bool(adap_mult))
    a.grid = int(grid)
    a.world = int(world)
    a.stats, a.grad_out = stats.ctypes.data, grad.ctypes.data
    lib().orc_ppo_update(C.byref(a))
    if ctx_sidx is not None:
        return stats, grad, ctx_loss
    return stats, grad


SAMPLER_IDS = {"l2": 0, "unit_square": 1, "positive_square": 2, "categorical": 3, "natural_numbers": 4}


def adap_draw(n, K, C, sampler, seed, stream, index0=0, S=0, n_mb=1, M=0, batch_size=0):
    """Philox stand-in for the draws of get_context_kl_loss: (sidx [n, S] or None, draws [n, K, C])."""
    sidx = np.zeros((n, S), np.int32) if S > 0 else None
    draws = np.zeros((n, K, C), np.float32)
    lib().orc_adap_draw(None if sidx is None else sidx.ctypes.data_as(C_.c_void_p), draws.ctypes.data_as(C_.c_void_p),
                        C_.c_int64(n), C_.c_int64(n_mb), C_.c_int64(M), C_.c_int64(batch_size), C_.c_int32(S),
                        C_.c_int32(K), C_.c_int32(C), C_.c_int32(SAMPLER_IDS[sampler]), C_.c_uint64(seed),
                        C_.c_uint32(stream), C_.c_uint32(index0))
    return sidx, draws


class OrcModularArgs(C.Structure):
    _fields_ = [
        ("space", C.POINTER(OrcSpace)),
        ("params", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
        ("adam_step", C.c_int64), ("vf_step", C.c_int64),
        ("num_partners", C.c_int32), ("partner_idx", C.c_int32),
        ("obs", C.c_void_p), ("actions", C.c_void_p), ("old_logp", C.c_void_p),
        ("advantages", C.c_void_p), ("returns", C.c_void_p), ("index", C.c_void_p), ("perm", C.c_void_p),
        ("M", C.c_int64), ("batch_size", C.c_int64), ("n_epochs", C.c_int32),
        ("learning_rate", C.c_float), ("clip_range", C.c_float), ("ent_coef", C.c_float),
        ("vf_coef", C.c_float), ("max_grad_norm", C.c_float),
        ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("adam_eps", C.c_float),
        ("marginal_reg_coef", C.c_float), ("grid", C.c_int32),
        ("stats", C.c_void_p), ("marginal_out", C.c_void_p),
    ]


def modular_param_count(space, num_partners):
    lib().orc_modular_param_count.restype = C.c_int64
    return int(lib().orc_modular_param_count(C.byref(space), C.c_int32(num_partners)))


def modular_update(space, params, adam_m, adam_v, adam_step, vf_step, num_partners, partner_idx, obs, actions, old_logp,
                   advantages, returns, perm, batch_size, grid, index=None, learning_rate=3e-4, clip_range=0.2,
                   ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, betas=(0.9, 0.999), eps=1e-5, marginal_reg_coef=0.0):
    """One partner's phase of ModularAlgorithm.train; params / adam state are updated IN PLACE.
    Returns (stats [n_epochs * n_mb, 8], marginal [n_epochs * n_mb])."""
    f32 = lambda x: np.ascontiguousarray(x, np.float32)  # noqa: E731
    for a_ in (params, adam_m, adam_v):
        assert a_.dtype == np.float32 and a_.flags.c_contiguous
    if space.obs_kind == 1:
        obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 64)
    else:
        obs = np.ascontiguousarray(obs, np.uint8).reshape(-1, 32 if space.obs_len <= 32 else 96)
    actions = np.ascontiguousarray(actions, np.uint8).reshape(-1, 4)
    old_logp, advantages, returns = f32(old_logp).reshape(-1), f32(advantages).reshape(-1), f32(returns).reshape(-1)
    perm = np.ascontiguousarray(perm, np.int32)
    n_epochs, M = perm.shape
    n_mb = (M + batch_size - 1) // batch_size
    stats = np.zeros((n_epochs * n_mb, 8), np.float32)
    marg = np.zeros(n_epochs * n_mb, np.float32)
    a = OrcModularArgs()
    a.space = C.pointer(space)
    a.params, a.adam_m, a.adam_v = params.ctypes.data, adam_m.ctypes.data, adam_v.ctypes.data
    a.adam_step, a.vf_step = int(adam_step), int(vf_step)
    a.num_partners, a.partner_idx = int(num_partners), int(partner_idx)
    a.obs, a.actions = obs.ctypes.data, actions.ctypes.data
    a.old_logp, a.advantages, a.returns = old_logp.ctypes.data, advantages.ctypes.data, returns.ctypes.data
    if index is not None:
        index = np.ascontiguousarray(index, np.int32)
        a.index = index.ctypes.data
    a.perm = perm.ctypes.data
    a.M, a.batch_size, a.n_epochs = M, int(batch_size), n_epochs
    a.learning_rate, a.clip_range, a.ent_coef = learning_rate, clip_range, ent_coef
    a.vf_coef, a.max_grad_norm = vf_coef, max_grad_norm
    a.adam_beta1, a.adam_beta2, a.adam_eps = betas[0], betas[1], eps
    a.marginal_reg_coef, a.grid = float(marginal_reg_coef), int(grid)
    a.stats, a.marginal_out = stats.ctypes.data, marg.ctypes.data
    lib().orc_modular_update(C.byref(a))
    return stats, marg


def modular_forward(space, params, num_partners, partner_idx, obs, seed=0, rng_stream=2, tick=0, slot=0, idx0=0,
                    action_in=None):
    params = np.ascontiguousarray(params, np.float32)
    obs = np.ascontiguousarray(obs, np.uint8) if space.obs_kind == 0 else np.ascontiguousarray(obs, np.float32)
    B, stride = obs.shape
    L = sum(space.head_n[i] for i in range(space.n_heads))
    action = np.zeros((B, 4), np.uint8)
    value, logp, ent = np.empty(B, np.float32), np.empty(B, np.float32), np.empty(B, np.float32)
    logits = np.empty((B, L), np.float32)
    if action_in is not None:
        action_in = np.ascontiguousarray(action_in, np.uint8)
    p_ = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib().orc_modular_forward(C.byref(space), p_(params), C.c_int32(num_partners), C.c_int32(partner_idx), p_(obs),
                              C.c_int64(stride), C.c_int64(B), C.c_uint64(seed), C.c_uint32(rng_stream),
                              C.c_uint32(tick), C.c_uint32(slot), C.c_int64(idx0), p_(action_in), p_(action),
                              p_(value), p_(logp), p_(ent), p_(logits))
    return dict(action=action, value=value, logp=logp, entropy=ent, logits=logits)
