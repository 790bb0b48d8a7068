"""Host-side helpers of the one-partner-per-GPU sharding; pure
index arithmetic, usable on CPU tensors (the gloo tests) and CUDA tensors."""
import torch

PACKED_BYTES = 48
# byte offsets inside one packed ego-transition record
OFF_OBS, OFF_ACTION, OFF_LOGP, OFF_ADV, OFF_RET = 0, 32, 36, 40, 44


def global_env_major_index(world, T, N, device="cpu"):
    """Sample order of the all-gathered ego stream: rank r's record (t, n) sits at
    r*T*N + t*N + n; samples are enumerated env-major over GLOBAL env ids
    g = r*N + n (SB3 swap_and_flatten): j = g*T + t."""
    r = torch.arange(world, device=device, dtype=torch.int64).view(-1, 1, 1)
    n = torch.arange(N, device=device, dtype=torch.int64).view(1, -1, 1)
    t = torch.arange(T, device=device, dtype=torch.int64).view(1, 1, -1)
    return (r * T * N + t * N + n).reshape(-1).to(torch.int32).contiguous()


def shard_env0(rank, n_envs_per_rank):
    """First global env id of a rank's shard (indexes the global Philox streams)."""
    return rank * n_envs_per_rank
