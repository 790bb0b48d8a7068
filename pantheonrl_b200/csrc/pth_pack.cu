// pth_pack.cu — multi-GPU exchange: the ego's rollout transitions are packed
// into records (obs row | action 4 | old_logp 4 | advantage 4 | return 4: 48 bytes for one-hot
// observation rows of 32 bytes, 272 bytes for Box rows of 64 floats) so that ONE all-gather per
// rollout makes the whole ego batch resident on every GPU; the update kernel then reads the
// packed stream in place (rec_stride = record bytes).
//
//   pth_pack_transitions     : pack into a local staging buffer
//   pth_allgather_transitions: ncclAllGather of the staging buffers (the library's own NCCL
//                              communicator: pth_comm_unique_id / pth_comm_init)
//   pth_pack_allgather_p2p   : pack and store each record straight into EVERY rank's gather buffer
//                              through NVLink peer mappings — pack and all-gather are one kernel.
//
// Work item = one 16-byte chunk of one record (obs_bytes / 16 observation chunks + 1 tail chunk):
// consecutive threads read consecutive 16-byte pieces of the observation rows and write
// consecutive pieces of the stream, for both record sizes.
#include <dlfcn.h>

#include "pth_common.cuh"

namespace {

__device__ __forceinline__ uint4 load_chunk(const uint8_t* obs, const uint8_t* act, const float* logp,
                                            const float* adv, const float* ret, int64_t i, int j, int oc) {
  if (j < oc) return __ldcs(reinterpret_cast<const uint4*>(obs + i * (int64_t)(oc * 16)) + j);
  uint4 r;
  r.x = __ldcs(reinterpret_cast<const uint32_t*>(act) + i);
  r.y = __float_as_uint(__ldcs(logp + i));
  r.z = __float_as_uint(__ldcs(adv + i));
  r.w = __float_as_uint(__ldcs(ret + i));
  return r;
}

__global__ void pack_kernel(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ act,
                            const float* __restrict__ logp, const float* __restrict__ adv,
                            const float* __restrict__ ret, int64_t count, int oc, uint8_t* __restrict__ out) {
  const int rc = oc + 1;  // chunks per record
  const int64_t total = count * rc;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / rc;
    const int j = (int)(q - i * rc);
    reinterpret_cast<uint4*>(out)[q] = load_chunk(obs, act, logp, adv, ret, i, j, oc);
  }
}

__global__ void pack_p2p_kernel(const uint8_t* __restrict__ obs, const uint8_t* __restrict__ act,
                                const float* __restrict__ logp, const float* __restrict__ adv,
                                const float* __restrict__ ret, int64_t count, int oc,
                                uint8_t* const* __restrict__ peers, int world, int rank) {
  const int rc = oc + 1;
  const int64_t total = count * rc;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / rc;
    const int j = (int)(q - i * rc);
    const uint4 v = load_chunk(obs, act, logp, adv, ret, i, j, oc);
    // own copy first, then the peers in ring order so that the NVLink egress of
    // all ranks is spread over all links at any moment
    for (int k = 0; k < world; ++k) {
      const int dst = (rank + k) % world;
      reinterpret_cast<uint4*>(peers[dst])[(int64_t)rank * total + q] = v;
    }
  }
}

int pack_grid(const pth_ctx* ctx, int64_t chunks) {
  int grid = pth_ceil_div(chunks, 256);
  const int cap = ctx->sm_count * 8;
  return grid > cap ? cap : grid;
}

// ---- NCCL, bound at run time (dlopen) so that the library loads on hosts without it
typedef struct { char internal[128]; } nccl_uid;  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm;
struct NcclApi {
  void* handle;
  int (*GetUniqueId)(nccl_uid*);
  int (*CommInitRank)(nccl_comm*, int, nccl_uid, int);
  int (*CommDestroy)(nccl_comm);
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
  const char* (*GetErrorString)(int);
};

NcclApi* nccl_api() {
  static NcclApi api = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  static int tried = 0;
  if (!tried) {
    tried = 1;
    // a process that already holds libnccl.so.2 (e.g. PyTorch's copy) gets that one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (int (*)(nccl_uid*))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(nccl_comm*, int, nccl_uid, int))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (int (*)(nccl_comm))dlsym(h, "ncclCommDestroy");
      api.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm, cudaStream_t))dlsym(h, "ncclAllGather");
      api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather) api.handle = h;
    }
  }
  return api.handle ? &api : nullptr;
}

#define PTH_NCCL(api, expr)                                                                   \
  do {                                                                                        \
    const int rc__ = (expr);                                                                  \
    if (rc__ != 0) {                                                                          \
      pth_set_error("%s failed: NCCL error %d (%s)", #expr, rc__,                            \
                    (api)->GetErrorString ? (api)->GetErrorString(rc__) : "?");               \
      return PTH_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

}  // namespace

extern "C" int pth_pack_transitions(pth_ctx* ctx, const uint8_t* d_obs, int32_t obs_bytes,
                                    const uint8_t* d_actions, const float* d_logp,
                                    const float* d_advantages, const float* d_returns, int64_t count,
                                    uint8_t* d_packed, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_obs && d_actions && d_logp && d_advantages && d_returns && d_packed,
                "NULL device pointer");
  PTH_CHECK_ARG(count >= 0, "negative count");
  PTH_CHECK_ARG(obs_bytes > 0 && obs_bytes % 16 == 0, "observation rows must be a multiple of 16 bytes");
  PTH_CHECK_ARG(((uintptr_t)d_obs % 16) == 0 && ((uintptr_t)d_packed % 16) == 0,
                "obs / packed must be 16-byte aligned");
  if (count == 0) return PTH_OK;
  const int oc = obs_bytes / 16;
  pack_kernel<<<pack_grid(ctx, count * (oc + 1)), 256, 0, (cudaStream_t)stream>>>(
      d_obs, d_actions, d_logp, d_advantages, d_returns, count, oc, d_packed);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_pack_allgather_p2p(pth_ctx* ctx, const uint8_t* d_obs, int32_t obs_bytes,
                                      const uint8_t* d_actions, const float* d_logp,
                                      const float* d_advantages, const float* d_returns, int64_t count,
                                      uint8_t* const* d_peer_bufs, int32_t world, int32_t rank,
                                      void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_obs && d_actions && d_logp && d_advantages && d_returns && d_peer_bufs,
                "NULL device pointer");
  PTH_CHECK_ARG(count >= 0 && world >= 1 && rank >= 0 && rank < world, "bad count / world / rank");
  PTH_CHECK_ARG(obs_bytes > 0 && obs_bytes % 16 == 0, "observation rows must be a multiple of 16 bytes");
  PTH_CHECK_ARG(((uintptr_t)d_obs % 16) == 0, "obs must be 16-byte aligned");
  if (count == 0) return PTH_OK;
  const int oc = obs_bytes / 16;
  pack_p2p_kernel<<<pack_grid(ctx, count * (oc + 1)), 256, 0, (cudaStream_t)stream>>>(
      d_obs, d_actions, d_logp, d_advantages, d_returns, count, oc, d_peer_bufs, world, rank);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_comm_unique_id(void* out128) {
  PTH_CHECK_ARG(out128 != nullptr, "NULL output");
  NcclApi* api = nccl_api();
  if (!api) {
    pth_set_error("pth_comm_unique_id: libnccl.so.2 not found");
    return PTH_ENOSUP;
  }
  PTH_NCCL(api, api->GetUniqueId(reinterpret_cast<nccl_uid*>(out128)));
  return PTH_OK;
}

extern "C" int pth_comm_init(pth_ctx* ctx, const void* unique_id128, int32_t world, int32_t rank) {
  PTH_CHECK_ARG(ctx != nullptr && unique_id128 != nullptr, "NULL ctx / id");
  PTH_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad world / rank");
  PTH_CHECK_ARG(ctx->nccl_comm == nullptr, "this context already has a communicator");
  NcclApi* api = nccl_api();
  if (!api) {
    pth_set_error("pth_comm_init: libnccl.so.2 not found");
    return PTH_ENOSUP;
  }
  PTH_CUDA(cudaSetDevice(ctx->device));
  nccl_uid id;
  memcpy(&id, unique_id128, sizeof(id));
  nccl_comm comm = nullptr;
  PTH_NCCL(api, api->CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
  ctx->nccl_world = world;
  ctx->nccl_rank = rank;
  return PTH_OK;
}

extern "C" int pth_comm_destroy(pth_ctx* ctx) {
  PTH_CHECK_ARG(ctx != nullptr, "NULL ctx");
  if (ctx->nccl_comm) {
    NcclApi* api = nccl_api();
    if (api) api->CommDestroy((nccl_comm)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return PTH_OK;
}

extern "C" int pth_allgather_transitions(pth_ctx* ctx, const uint8_t* d_packed, int64_t bytes_per_rank,
                                         uint8_t* d_gathered, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && d_packed && d_gathered, "NULL pointer");
  PTH_CHECK_ARG(bytes_per_rank >= 0, "negative size");
  PTH_CHECK_ARG(ctx->nccl_comm != nullptr, "no communicator: call pth_comm_init first");
  NcclApi* api = nccl_api();
  PTH_CHECK_ARG(api != nullptr, "libnccl.so.2 not found");
  if (bytes_per_rank == 0) return PTH_OK;
  PTH_NCCL(api, api->AllGather(d_packed, d_gathered, (size_t)bytes_per_rank, /*ncclUint8*/ 1,
                               (nccl_comm)ctx->nccl_comm, (cudaStream_t)stream));
  return PTH_OK;
}
