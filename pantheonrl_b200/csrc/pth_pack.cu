// pth_pack.cu — multi-GPU exchange staging: the ego's rollout
// transitions are packed into 48-byte records (obs 32 | action 4 | old_logp 4 |
// advantage 4 | return 4) so that ONE all-gather per rollout makes the whole
// ego batch resident on every GPU; the update kernel then reads the packed
// stream in place (rec_stride = 48).
//
//   pth_pack_transitions   : pack into a local staging buffer (NCCL all-gather follows)
//   pth_pack_allgather_p2p : pack and store each record straight into EVERY
//                            rank's gather buffer through NVLink peer mappings —
//                            the pack and the all-gather are one kernel, no staging.
#include "pth_common.cuh"

namespace {

struct Rec {
  uint4 a, b, c;  // 48 bytes
};

__device__ __forceinline__ Rec load_rec(const uint8_t* obs, const uint8_t* act, const float* logp,
                                        const float* adv, const float* ret, int64_t i) {
  const uint4* o = reinterpret_cast<const uint4*>(obs + i * 32);
  Rec r;
  r.a = __ldcs(o);
  r.b = __ldcs(o + 1);
  r.c.x = __ldcs(reinterpret_cast<const uint32_t*>(act) + i);
  r.c.y = __float_as_uint(__ldcs(logp + i));
  r.c.z = __float_as_uint(__ldcs(adv + i));
  r.c.w = __float_as_uint(__ldcs(ret + i));
  return r;
}

__device__ __forceinline__ void store_rec(uint8_t* dst, int64_t i, const Rec& r) {
  uint4* q = reinterpret_cast<uint4*>(dst + i * PTH_PACKED_BYTES);
  q[0] = r.a;
  q[1] = r.b;
  q[2] = r.c;
}

__global__ void pack_kernel(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ act,
                            const float* __restrict__ logp, const float* __restrict__ adv,
                            const float* __restrict__ ret, int64_t count, uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    store_rec(out, i, load_rec(obs, act, logp, adv, ret, i));
}

__global__ void pack_p2p_kernel(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ act,
                                const float* __restrict__ logp, const float* __restrict__ adv,
                                const float* __restrict__ ret, int64_t count,
                                uint8_t* const* __restrict__ peers, int world, int rank) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const Rec r = load_rec(obs, act, logp, adv, ret, i);
    // own copy first, then the peers in ring order so that the NVLink egress of
    // all ranks is spread over all links at any moment
    for (int k = 0; k < world; ++k) {
      const int dst = (rank + k) % world;
      store_rec(peers[dst], (int64_t)rank * count + i, r);
    }
  }
}

}  // namespace

extern "C" int pth_pack_transitions(pth_ctx* ctx, const uint8_t* d_obs, const uint8_t* d_actions,
                                    const float* d_logp, const float* d_advantages,
                                    const float* d_returns, int64_t count, uint8_t* d_packed,
                                    void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_obs && d_actions && d_logp && d_advantages && d_returns && d_packed,
                "NULL device pointer");
  PTH_CHECK_ARG(count >= 0, "negative count");
  PTH_CHECK_ARG(((uintptr_t)d_obs % 16) == 0 && ((uintptr_t)d_packed % 16) == 0,
                "obs / packed must be 16-byte aligned");
  if (count == 0) return PTH_OK;
  int grid = pth_ceil_div(count, 256);
  const int cap = ctx->sm_count * 8;
  if (grid > cap) grid = cap;
  pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_obs, d_actions, d_logp, d_advantages,
                                                      d_returns, count, d_packed);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_pack_allgather_p2p(pth_ctx* ctx, const uint8_t* d_obs, const uint8_t* d_actions,
                                      const float* d_logp, const float* d_advantages,
                                      const float* d_returns, int64_t count,
                                      uint8_t* const* d_peer_bufs, int32_t world, int32_t rank,
                                      void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_obs && d_actions && d_logp && d_advantages && d_returns && d_peer_bufs,
                "NULL device pointer");
  PTH_CHECK_ARG(count >= 0 && world >= 1 && rank >= 0 && rank < world, "bad count / world / rank");
  PTH_CHECK_ARG(((uintptr_t)d_obs % 16) == 0, "obs must be 16-byte aligned");
  if (count == 0) return PTH_OK;
  int grid = pth_ceil_div(count, 256);
  const int cap = ctx->sm_count * 8;
  if (grid > cap) grid = cap;
  pack_p2p_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_obs, d_actions, d_logp, d_advantages,
                                                          d_returns, count, d_peer_bufs, world, rank);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}
