"""ctypes binding of libpantheon_b200.so (include/pantheon_b200.h).

There is no fallback: if the shared library is missing, or a call returns an
error, this module raises.  PyTorch tensors are passed as raw device pointers
(``tensor.data_ptr()``) plus the current CUDA stream handle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpantheon_b200.so")

PTH_MAX_OBS_SLOTS = 96
PTH_MAX_HEADS = 4
PTH_HIDDEN = 64
PTH_OBS_ONEHOT, PTH_OBS_BOX = 0, 1
PTH_ENV_RPS, PTH_ENV_LIAR, PTH_ENV_OVERCOOKED = 0, 1, 2
PTH_OC_MAX_CELLS, PTH_OC_MAX_POTS, PTH_OC_OBS, PTH_OC_ROW = 128, 4, 62, 64
PTH_OC_STATE_BYTES = 40
PTH_OC_FLOOR, PTH_OC_COUNTER, PTH_OC_ONION, PTH_OC_POT, PTH_OC_DISH, PTH_OC_SERVE = range(6)
PTH_LOSS_PPO, PTH_LOSS_BC, PTH_LOSS_ADAP, PTH_LOSS_MODULAR = 0, 1, 2, 3
PTH_UPDATE_FLAG_WORDS = 16384  # include/pantheon_b200.h
PTH_PACKED_BYTES, PTH_PACKED_BYTES_BOX = 48, 272

STREAM_ENV, STREAM_EGO, STREAM_ALT, STREAM_SHUFFLE_EGO, STREAM_SHUFFLE_ALT = 1, 2, 3, 4, 5


class PthError(RuntimeError):
    pass


class Space(C.Structure):
    _fields_ = [
        ("obs_kind", C.c_int32),
        ("obs_len", C.c_int32),
        ("obs_nvec", C.c_int32 * PTH_MAX_OBS_SLOTS),
        ("n_heads", C.c_int32),
        ("head_n", C.c_int32 * PTH_MAX_HEADS),
    ]

    @classmethod
    def onehot(cls, nvec, heads):
        s = cls()
        if len(nvec) > len(s.obs_nvec) or len(heads) > len(s.head_n):
            raise PthError(f"one-hot spaces take at most {len(s.obs_nvec)} observation slots and "
                           f"{len(s.head_n)} action heads (got {len(nvec)} / {len(heads)})")
        s.obs_kind = PTH_OBS_ONEHOT
        s.obs_len = len(nvec)
        for i, v in enumerate(nvec):
            s.obs_nvec[i] = int(v)
        s.n_heads = len(heads)
        for i, v in enumerate(heads):
            s.head_n[i] = int(v)
        return s

    @classmethod
    def box(cls, dim, heads):
        s = cls()
        s.obs_kind = PTH_OBS_BOX
        s.obs_len = int(dim)
        s.n_heads = len(heads)
        for i, v in enumerate(heads):
            s.head_n[i] = int(v)
        return s

    @property
    def row_bytes(self):
        """Bytes of one observation row: PTH_OBS_ROW_BYTES for one-hot spaces, 64 fp32 for Box spaces."""
        return 4 * PTH_OC_ROW if self.obs_kind == PTH_OBS_BOX else (32 if self.obs_len <= 32 else 96)

    @property
    def nvec(self):
        return [self.obs_nvec[i] for i in range(self.obs_len)]

    @property
    def heads(self):
        return [self.head_n[i] for i in range(self.n_heads)]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("space", C.POINTER(Space)),
        ("d_params", C.c_void_p),
        ("d_obs", C.c_void_p),
        ("obs_stride", C.c_int64),
        ("B", C.c_int64),
        ("seed", C.c_uint64),
        ("rng_stream", C.c_uint32),
        ("tick", C.c_uint32),
        ("slot", C.c_uint32),
        ("idx0", C.c_int64),
        ("d_action_in", C.c_void_p),
        ("d_action", C.c_void_p),
        ("d_value", C.c_void_p),
        ("d_logp", C.c_void_p),
        ("d_entropy", C.c_void_p),
        ("d_logits", C.c_void_p),
        ("d_race", C.c_void_p),
        ("context_size", C.c_int32),
        ("d_context", C.c_void_p),
        ("context_stride", C.c_int64),
        ("num_partners", C.c_int32),
        ("partner_idx", C.c_int32),
        ("adap_mult", C.c_int32),
    ]


class Buffer(C.Structure):
    _fields_ = [
        ("d_obs", C.c_void_p),
        ("d_actions", C.c_void_p),
        ("d_rewards", C.c_void_p),
        ("d_values", C.c_void_p),
        ("d_logp", C.c_void_p),
        ("d_episode_starts", C.c_void_p),
        ("d_count", C.c_void_p),
        ("Tcap", C.c_int64),
    ]


class EnvCarry(C.Structure):
    _fields_ = [
        ("d_ego_last_start", C.c_void_p),
        ("d_alt_last_done", C.c_void_p),
        ("d_total_rew", C.c_void_p),
        ("d_flags", C.c_void_p),
        ("d_game_state", C.c_void_p),
        ("d_ego_last_value", C.c_void_p),
        ("d_ego_last_done", C.c_void_p),
        ("d_ep_stats", C.c_void_p),
        ("d_alt_boot_done", C.c_void_p),
    ]


class OvercookedLayout(C.Structure):
    """pth_overcooked_layout (include/pantheon_b200.h): inputs + tables derived by
    pth_overcooked_layout_init."""
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("cook_time", C.c_int32), ("num_items", C.c_int32), ("delivery_reward", C.c_int32),
        ("horizon", C.c_int32),
        ("rew_placement_in_pot", C.c_int32), ("rew_dish_pickup", C.c_int32), ("rew_soup_pickup", C.c_int32),
        ("ego_agent_idx", C.c_int32),
        ("start_x", C.c_int32 * 2), ("start_y", C.c_int32 * 2),
        ("terrain", C.c_uint8 * PTH_OC_MAX_CELLS),
        ("n_pots", C.c_int32), ("n_counters", C.c_int32),
        ("slot", C.c_uint8 * PTH_OC_MAX_CELLS),
        ("wall", C.c_uint8 * PTH_OC_MAX_CELLS),
        ("pot_x", C.c_uint8 * PTH_OC_MAX_POTS), ("pot_y", C.c_uint8 * PTH_OC_MAX_POTS),
        ("static_delta", C.c_int8 * (PTH_OC_MAX_CELLS * 4 * 3 * 2)),
        ("pot_dist", C.c_uint8 * (PTH_OC_MAX_CELLS * 4 * PTH_OC_MAX_POTS)),
    ]


class RolloutArgs(C.Structure):
    _fields_ = [
        ("env_kind", C.c_int32),
        ("partner_records", C.c_int32),
        ("space", C.POINTER(Space)),
        ("d_ego_params", C.c_void_p),
        ("d_alt_params", C.c_void_p),
        ("ego", Buffer),
        ("alt", Buffer),
        ("carry", EnvCarry),
        ("N", C.c_int64),
        ("T", C.c_int64),
        ("env0", C.c_int64),
        ("seed", C.c_uint64),
        ("tick0", C.c_uint32),
        ("probegostart", C.c_float),
        ("first_rollout", C.c_int32),
        ("d_layout", C.c_void_p),
    ]


class UpdateArgs(C.Structure):
    _fields_ = [
        ("space", C.POINTER(Space)),
        ("d_params", C.c_void_p),
        ("d_adam_m", C.c_void_p),
        ("d_adam_v", C.c_void_p),
        ("adam_step", C.c_int64),
        ("d_obs", C.c_void_p),
        ("d_obs_f32", C.c_void_p),
        ("obs_stride", C.c_int64),
        ("d_actions", C.c_void_p),
        ("d_old_logp", C.c_void_p),
        ("d_advantages", C.c_void_p),
        ("d_returns", C.c_void_p),
        ("rec_stride", C.c_int64),
        ("d_index", C.c_void_p),
        ("d_perm", C.c_void_p),
        ("M", C.c_int64),
        ("batch_size", C.c_int64),
        ("n_epochs", C.c_int32),
        ("learning_rate", C.c_float),
        ("clip_range", C.c_float),
        ("ent_coef", C.c_float),
        ("vf_coef", C.c_float),
        ("max_grad_norm", C.c_float),
        ("adam_beta1", C.c_float),
        ("adam_beta2", C.c_float),
        ("adam_eps", C.c_float),
        ("normalize_advantage", C.c_int32),
        ("grid_ctas", C.c_int32),
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("peer_xbuf", C.POINTER(C.c_void_p)),
        ("peer_flags", C.POINTER(C.c_void_p)),
        ("flag_epoch", C.c_uint32),
        ("d_workspace", C.c_void_p),
        ("workspace_bytes", C.c_int64),
        ("d_stats", C.c_void_p),
        ("loss_kind", C.c_int32),
        ("l2_weight", C.c_float),
        ("context_size", C.c_int32),
        ("d_context", C.c_void_p),
        ("context_loss_coeff", C.c_float),
        ("num_context_samples", C.c_int32),
        ("num_state_samples", C.c_int32),
        ("d_ctx_states", C.c_void_p),
        ("d_ctx_draws", C.c_void_p),
        ("d_ctx_loss", C.c_void_p),
        ("num_partners", C.c_int32),
        ("partner_idx", C.c_int32),
        ("partner_vf_step", C.c_int64),
        ("marginal_reg_coef", C.c_float),
        ("d_modular_scratch", C.c_void_p),
        ("modular_scratch_bytes", C.c_int64),
        ("adap_mult", C.c_int32),
    ]


_vp, _i64, _i32, _u64, _u32, _f, _d = (C.c_void_p, C.c_int64, C.c_int32, C.c_uint64,
                                       C.c_uint32, C.c_float, C.c_double)

# name -> (restype, argtypes).  Every symbol declared in include/pantheon_b200.h.
SIGNATURES = {
    "pth_version": (C.c_int, []),
    "pth_last_error": (C.c_char_p, []),
    "pth_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "pth_ctx_destroy": (C.c_int, [_vp]),
    "pth_ctx_sm_count": (C.c_int, [_vp]),
    "pth_sync_debug": (C.c_int, [_vp, _vp]),
    "pth_space_feature_dim": (C.c_int, [C.POINTER(Space)]),
    "pth_space_logit_dim": (C.c_int, [C.POINTER(Space)]),
    "pth_policy_param_count": (_i64, [C.POINTER(Space)]),
    "pth_gae_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _d, _d, C.c_int, _vp]),
    "pth_gae_ragged_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _d, _d, _vp]),
    "pth_env_rps_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pth_env_liar_reset": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _u64, _u32, _i64, _f, _vp]),
    "pth_env_liar_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pth_overcooked_layout_init": (C.c_int, [C.POINTER(OvercookedLayout)]),
    "pth_env_overcooked_reset": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "pth_env_overcooked_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pth_policy_forward": (C.c_int, [_vp, C.POINTER(ForwardArgs), _vp]),
    "pth_debug_math": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _vp]),
    "pth_debug_ffma_peak": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "pth_rollout_run": (C.c_int, [_vp, C.POINTER(RolloutArgs), _vp]),
    "pth_perm_feistel": (C.c_int, [_vp, _vp, _i64, _i32, _u64, _u32, _u32, _vp]),
    "pth_index_build": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "pth_update_grid": (C.c_int, [_vp, C.POINTER(Space), _i64, _i64]),
    "pth_update_xbuf_bytes": (_i64, [C.POINTER(Space), _i32]),
    "pth_index_workspace_bytes": (_i64, [_i64]),
    "pth_update_workspace_bytes": (_i64, [_vp, C.POINTER(Space), _i64, _i64]),
    "pth_adap_workspace_bytes": (_i64, [_vp, C.POINTER(Space), C.c_int32, _i64, _i64]),
    "pth_adap_param_count": (_i64, [C.POINTER(Space), C.c_int32]),
    "pth_modular_param_count": (_i64, [C.POINTER(Space), C.c_int32]),
    "pth_adap_mult_param_count": (_i64, [C.POINTER(Space), C.c_int32]),
    "pth_adap_mult_workspace_bytes": (_i64, [_vp, C.POINTER(Space), C.c_int32, _i64, _i64]),
    "pth_adap_mult_scratch_bytes": (_i64, [_vp, C.POINTER(Space), C.c_int32]),
    "pth_modular_workspace_bytes": (_i64, [_vp, C.POINTER(Space), C.c_int32, _i64, _i64]),
    "pth_modular_scratch_bytes": (_i64, [_vp, C.POINTER(Space), C.c_int32]),
    "pth_adap_draw": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_uint64, C.c_uint32, C.c_uint32, _vp]),
    "pth_ppo_update": (C.c_int, [_vp, C.POINTER(UpdateArgs), _vp]),
    "pth_debug_update_profile": (C.c_int, [_vp]),
    "pth_pack_transitions": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "pth_pack_allgather_p2p": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp]),
    "pth_comm_unique_id": (C.c_int, [_vp]),
    "pth_comm_init": (C.c_int, [_vp, _vp, _i32, _i32]),
    "pth_comm_destroy": (C.c_int, [_vp]),
    "pth_allgather_transitions": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
}

_lib = None
LAUNCHES = 0  # kernels of libpantheon_b200.so launched through the bindings (bench.py reports it)


def count_launch(n=1):
    global LAUNCHES
    LAUNCHES += n


def load():
    """dlopen the library and bind every declared symbol.  No device calls."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PthError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().pth_last_error().decode("utf-8", "replace")
        raise PthError(f"{what} failed ({rc}): {msg}")


class Context:
    """pth_ctx wrapper; one per (process, device)."""

    _cache = {}

    def __init__(self, device=0):
        lib = load()
        h = _vp()
        check(lib.pth_ctx_create(int(device), C.byref(h)), "pth_ctx_create")
        self.handle = h
        self.device = int(device)
        self.sm_count = lib.pth_ctx_sm_count(h)
        self.has_comm = False  # pth_comm_init done on this context

    @classmethod
    def get(cls, device=0):
        device = int(device)
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]


def ptr(t):
    """Device (or host) pointer of a tensor, None -> NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
