// pth_envs.cu — standalone vectorised env step / reset kernels (one thread per
// env).  Used by the Python facade when the caller drives envs step by step and
// by the parity tests that replay the reference's own env traces.
#include "pth_games.cuh"

namespace {

__global__ void rps_step_kernel(const int32_t* __restrict__ ego_a, const int32_t* __restrict__ alt_a,
                                float* __restrict__ r_ego, float* __restrict__ r_alt, int64_t N) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float re, ra;
  pth_rps_outcome(ego_a[n], alt_a[n], re, ra);
  r_ego[n] = re;
  r_alt[n] = ra;
}

__device__ __forceinline__ void store_obs(uint8_t* obs, int64_t n, const uint32_t (&w)[8]) {
  uint4* q = reinterpret_cast<uint4*>(obs + n * 32);
  q[0] = make_uint4(w[0], w[1], w[2], w[3]);
  q[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

__global__ void liar_reset_kernel(pth_liar_state* __restrict__ state, uint8_t* __restrict__ ego_first,
                                  uint8_t* __restrict__ obs, int64_t N, uint64_t seed,
                                  uint32_t tick, int64_t env0, float probegostart) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  LiarRegs s;
  bool ef = liar_reset(s, seed, (uint64_t)(env0 + n), tick, 0u, probegostart);
  liar_store(state + n, s);
  if (ego_first) ego_first[n] = ef ? 1 : 0;
  if (obs) {
    uint32_t w[8];
    liar_obs(s, ef ? 0 : 1, w);
    store_obs(obs, n, w);
  }
}

__global__ void liar_step_kernel(pth_liar_state* __restrict__ state, const uint8_t* __restrict__ is_ego,
                                 const uint8_t* __restrict__ action, uint8_t* __restrict__ obs,
                                 float* __restrict__ r_ego, float* __restrict__ r_alt,
                                 uint8_t* __restrict__ done, int64_t N) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  LiarRegs s;
  liar_load(state + n, s);
  int player = is_ego[n] ? 0 : 1;
  float re, ra;
  bool d = liar_step(s, player, (int)action[2 * n], (int)action[2 * n + 1], re, ra);
  liar_store(state + n, s);
  uint32_t w[8];
  liar_obs(s, 1 - player, w);
  store_obs(obs, n, w);
  r_ego[n] = re;
  r_alt[n] = ra;
  done[n] = d ? 1 : 0;
}

}  // namespace

extern "C" int pth_env_rps_step(pth_ctx* ctx, const int32_t* d_ego_action,
                                const int32_t* d_alt_action, float* d_ego_reward,
                                float* d_alt_reward, int64_t N, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_ego_action && d_alt_action && d_ego_reward && d_alt_reward, "NULL device pointer");
  PTH_CHECK_ARG(N >= 0, "negative size");
  if (N == 0) return PTH_OK;
  rps_step_kernel<<<pth_ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(
      d_ego_action, d_alt_action, d_ego_reward, d_alt_reward, N);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_env_liar_reset(pth_ctx* ctx, pth_liar_state* d_state, uint8_t* d_ego_first,
                                  uint8_t* d_obs, int64_t N, uint64_t seed, uint32_t tick,
                                  int64_t env0, float probegostart, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_state != nullptr, "NULL state pointer");
  PTH_CHECK_ARG(((uintptr_t)d_state % 16) == 0 && (d_obs == nullptr || ((uintptr_t)d_obs % 16) == 0),
                "state/obs must be 16-byte aligned");
  PTH_CHECK_ARG(N >= 0, "negative size");
  if (N == 0) return PTH_OK;
  liar_reset_kernel<<<pth_ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(
      d_state, d_ego_first, d_obs, N, seed, tick, env0, probegostart);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_env_liar_step(pth_ctx* ctx, pth_liar_state* d_state, const uint8_t* d_is_ego,
                                 const uint8_t* d_action, uint8_t* d_obs, float* d_ego_reward,
                                 float* d_alt_reward, uint8_t* d_done, int64_t N, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_state && d_is_ego && d_action && d_obs && d_ego_reward && d_alt_reward && d_done,
                "NULL device pointer");
  PTH_CHECK_ARG(((uintptr_t)d_state % 16) == 0 && ((uintptr_t)d_obs % 16) == 0,
                "state/obs must be 16-byte aligned");
  PTH_CHECK_ARG(N >= 0, "negative size");
  if (N == 0) return PTH_OK;
  liar_step_kernel<<<pth_ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(
      d_state, d_is_ego, d_action, d_obs, d_ego_reward, d_alt_reward, d_done, N);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}
