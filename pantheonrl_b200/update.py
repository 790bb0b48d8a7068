"""Bindings of the update-side C-ABI entry points: pth_perm_feistel,
pth_index_build, pth_ppo_update (one cooperative launch = one PPO.train())."""
import ctypes as C

import torch

from . import _lib
from ._lib import Context, check, current_stream


def _ctx(t):
    return Context.get(t.device.index if t.device.index is not None else torch.cuda.current_device())


def perm_feistel(M, n_epochs, seed, stream_id, epoch0=0, device="cuda", out=None):
    """[n_epochs, M] int32 keyed permutations (stands in for np.random.permutation)."""
    perm = out if out is not None else torch.empty(n_epochs, M, dtype=torch.int32, device=device)
    check(_lib.load().pth_perm_feistel(_ctx(perm).handle, perm.data_ptr(), int(M), int(n_epochs),
                                       int(seed), int(stream_id), int(epoch0) & 0xffffffff,
                                       current_stream()), "pth_perm_feistel")
    _lib.count_launch()
    return perm


def index_build(count, T, N, device="cuda"):
    """Env-major list of the valid (row, env) cells: returns (index[T*N] int32, total[1] int32)."""
    lib = _lib.load()
    index = torch.empty(T * N, dtype=torch.int32, device=device)
    total = torch.zeros(1, dtype=torch.int32, device=device)
    ws = torch.empty(int(lib.pth_index_workspace_bytes(N)), dtype=torch.uint8, device=device)
    check(lib.pth_index_build(_ctx(index).handle, count.data_ptr() if count is not None else None,
                              int(T), int(N), index.data_ptr(), total.data_ptr(), ws.data_ptr(),
                              current_stream()), "pth_index_build")
    _lib.count_launch(2)
    return index, total


ADAP_SAMPLERS = {"l2": 0, "unit_square": 1, "positive_square": 2, "categorical": 3, "natural_numbers": 4}


def adap_draw(n, K, C_, sampler, seed, stream_id, index0=0, S=0, n_mb=1, M=0, batch_size=0, device="cuda"):
    """The random draws of ADAP on Philox (SAMPLERS / th.randperm of pantheonrl/algos/adap/util.py:42-106):
    (states int32 [n, S] or None, contexts float32 [n, K, C])."""
    draws = torch.empty(n, K, C_, dtype=torch.float32, device=device)
    states = torch.empty(n, S, dtype=torch.int32, device=device) if S > 0 else None
    check(_lib.load().pth_adap_draw(_ctx(draws).handle, states.data_ptr() if states is not None else None,
                                    draws.data_ptr(), int(n), int(n_mb), int(M), int(batch_size), int(S), int(K),
                                    int(C_), ADAP_SAMPLERS[sampler], int(seed), int(stream_id),
                                    int(index0) & 0xffffffff, current_stream()), "pth_adap_draw")
    _lib.count_launch()
    return states, draws


def update_grid(space, M, batch_size, device=0):
    return int(_lib.load().pth_update_grid(Context.get(device).handle, C.byref(space), int(M),
                                           int(batch_size)))


class UpdateWorkspace:
    """Caller-owned scratch for pth_ppo_update (the library never allocates)."""

    def __init__(self, space, M, batch_size, device="cuda", context_size=0, num_partners=0, adap_mult=False):
        ctx = Context.get(torch.device(device).index or 0)
        lib = _lib.load()
        if adap_mult:  # AdapPolicyMult: + the per-CTA activation scratch of both towers
            n = int(lib.pth_adap_mult_workspace_bytes(ctx.handle, C.byref(space), int(context_size), int(M), int(batch_size)))
            ns = int(lib.pth_adap_mult_scratch_bytes(ctx.handle, C.byref(space), int(context_size)))
            if ns <= 0:
                raise _lib.PthError("pth_adap_mult_scratch_bytes failed")
            self.modular_scratch = torch.empty(ns, dtype=torch.uint8, device=device)
        elif num_partners > 0:  # ModularPolicy: + the per-CTA activation scratch of every partner module
            n = int(lib.pth_modular_workspace_bytes(ctx.handle, C.byref(space), int(num_partners), int(M), int(batch_size)))
            ns = int(lib.pth_modular_scratch_bytes(ctx.handle, C.byref(space), int(num_partners)))
            if ns <= 0:
                raise _lib.PthError("pth_modular_scratch_bytes failed")
            self.modular_scratch = torch.empty(ns, dtype=torch.uint8, device=device)
        else:
            n = int(lib.pth_adap_workspace_bytes(ctx.handle, C.byref(space), int(context_size), int(M), int(batch_size)))
        if n <= 0:
            raise _lib.PthError("pth_update_workspace_bytes failed")
        self.buf = torch.empty(n, dtype=torch.uint8, device=device)
        self.M, self.batch_size = M, batch_size


def ppo_update(space, params, adam_m, adam_v, adam_step, obs, actions, old_logp, advantages,
               returns, perm, batch_size, workspace, index=None, M=None, rec_stride=0,
               learning_rate=3e-4, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5,
               betas=(0.9, 0.999), eps=1e-5, normalize_advantage=True, grid_ctas=0, stats=None,
               peers=None, loss_kind=0, l2_weight=0.0, context=None, context_loss_coeff=0.0, ctx_states=None,
               ctx_draws=None, ctx_loss=None, num_partners=0, partner_idx=0, partner_vf_step=0, marginal_reg_coef=0.0,
               adap_mult=False):
    """SB3 PPO.train() over flat sample arrays on the device, in place on
    params / adam_m / adam_v. Returns the stats tensor [n_epochs * n_mb, 8]."""
    n_epochs = perm.shape[0]
    M = int(M if M is not None else perm.shape[1])
    n_mb = (M + batch_size - 1) // batch_size
    if stats is None:
        stats = torch.zeros(n_epochs * n_mb, 8, dtype=torch.float32, device=params.device)
    a = _lib.UpdateArgs()
    a.space = C.pointer(space)
    a.d_params, a.d_adam_m, a.d_adam_v = params.data_ptr(), adam_m.data_ptr(), adam_v.data_ptr()
    a.adam_step = int(adam_step)
    if space.obs_kind == _lib.PTH_OBS_BOX:  # fp32 rows of 64
        a.d_obs, a.d_obs_f32, a.obs_stride = None, obs.data_ptr(), obs.shape[-1]
    else:
        a.d_obs, a.d_obs_f32, a.obs_stride = obs.data_ptr(), None, space.row_bytes
    a.d_actions = actions.data_ptr()
    a.d_old_logp, a.d_advantages, a.d_returns = old_logp.data_ptr(), advantages.data_ptr(), returns.data_ptr()
    a.rec_stride = int(rec_stride)
    a.d_index = index.data_ptr() if index is not None else None
    a.d_perm = perm.data_ptr()
    a.M, a.batch_size, a.n_epochs = M, int(batch_size), int(n_epochs)
    a.learning_rate, a.clip_range, a.ent_coef = learning_rate, clip_range, ent_coef
    a.vf_coef, a.max_grad_norm = vf_coef, max_grad_norm
    a.adam_beta1, a.adam_beta2, a.adam_eps = betas[0], betas[1], eps
    a.normalize_advantage = int(normalize_advantage)
    a.grid_ctas = int(grid_ctas)
    if peers is not None:  # sharded multi-GPU update: see PeerExchange
        a.world, a.rank = peers.world, peers.rank
        a.peer_xbuf, a.peer_flags = peers.xbuf_array, peers.flag_array
        a.flag_epoch = peers.epoch & 0xffffffff
        peers.epoch += n_epochs * n_mb
    a.d_workspace, a.workspace_bytes = workspace.buf.data_ptr(), workspace.buf.numel()
    a.d_stats = stats.data_ptr()
    a.loss_kind, a.l2_weight = int(loss_kind), float(l2_weight)
    if context is not None:  # AdapPolicy: float32 [rows, C] contexts stored with the samples
        if context.dtype != torch.float32 or not context.is_contiguous() or context.dim() != 2:
            raise ValueError("context must be a contiguous float32 [rows, C] tensor")
        a.context_size, a.d_context = context.shape[1], context.data_ptr()
    if ctx_states is not None:  # ADAP context loss: int32 [n_epochs * n_mb, S], float32 [n_epochs * n_mb, K, C]
        if ctx_states.dtype != torch.int32 or ctx_draws.dtype != torch.float32:
            raise ValueError("ctx_states must be int32, ctx_draws float32")
        if ctx_states.shape[0] != n_epochs * n_mb or ctx_draws.shape[0] != n_epochs * n_mb:
            raise ValueError("one row of ctx_states / ctx_draws per minibatch")
        a.context_loss_coeff = float(context_loss_coeff)
        a.num_state_samples, a.num_context_samples = ctx_states.shape[1], ctx_draws.shape[1]
        a.d_ctx_states, a.d_ctx_draws = ctx_states.data_ptr(), ctx_draws.data_ptr()
        if ctx_loss is not None:
            a.d_ctx_loss = ctx_loss.data_ptr()
    if adap_mult:  # AdapPolicyMult parameter layout (context= given); scratch from UpdateWorkspace(adap_mult=True)
        a.adap_mult = 1
        a.d_modular_scratch = workspace.modular_scratch.data_ptr()
        a.modular_scratch_bytes = workspace.modular_scratch.numel()
    if num_partners > 0:  # ModularAlgorithm.train, one partner phase (loss_kind PTH_LOSS_MODULAR)
        a.num_partners, a.partner_idx, a.partner_vf_step = int(num_partners), int(partner_idx), int(partner_vf_step)
        a.marginal_reg_coef = float(marginal_reg_coef)
        a.d_modular_scratch = workspace.modular_scratch.data_ptr()
        a.modular_scratch_bytes = workspace.modular_scratch.numel()
        if ctx_loss is not None:
            a.d_ctx_loss = ctx_loss.data_ptr()
    check(_lib.load().pth_ppo_update(_ctx(params).handle, C.byref(a), current_stream()), "pth_ppo_update")
    _lib.count_launch()
    return stats


class PeerExchange:
    """Symmetric-memory (NVLink peer-mapped) exchange buffers + flag words for the
    sharded multi-GPU update; one per process group, reused by every launch."""

    def __init__(self, space, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        nbytes = int(_lib.load().pth_update_xbuf_bytes(C.byref(space), self.world))
        self.xbuf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.flags = symm.empty(_lib.PTH_UPDATE_FLAG_WORDS, dtype=torch.int32, device=device)
        self.xbuf.zero_()
        self.flags.zero_()
        torch.cuda.synchronize()
        self.hx = symm.rendezvous(self.xbuf, group)
        self.hf = symm.rendezvous(self.flags, group)
        self.xbuf_array = (C.c_void_p * self.world)(*[int(x) for x in self.hx.buffer_ptrs])
        self.flag_array = (C.c_void_p * self.world)(*[int(x) for x in self.hf.buffer_ptrs])
        self.epoch = 0
        dist.barrier(group)
