// pth_gae.cu — GAE(lambda) advantages and returns over [T][N] rollout buffers.
//
// Reference: stable-baselines3 1.7.0 RolloutBuffer.compute_returns_and_advantage
// as called from pantheonrl/common/agents.py:127-130; the same recurrence is
// in-tree at overcookedgym/human_aware_rl/baselines/baselines/ppo2/runner.py:150-165.
//
// HBM-bound: 3 fp32 reads + 2 fp32 writes per (t, env) = 20 B.  The recurrence
// is sequential in t but independent per env, and its loads do not depend on
// the recurrence, so each thread keeps a window of U timesteps of loads in
// flight while it retires the previous window.
#include "pth_common.cuh"

namespace {

template <int VEC>
struct VecT;
template <>
struct VecT<1> {
  using type = float;
};
template <>
struct VecT<2> {
  using type = float2;
};
template <>
struct VecT<4> {
  using type = float4;
};

template <int VEC>
__device__ __forceinline__ void ld_stream(const float* p, float (&out)[VEC]) {
  using V = typename VecT<VEC>::type;
  V v = __ldcs(reinterpret_cast<const V*>(p));
  memcpy(out, &v, sizeof(V));
}
template <int VEC>
__device__ __forceinline__ void st_stream(float* p, const float (&in)[VEC]) {
  using V = typename VecT<VEC>::type;
  V v;
  memcpy(&v, in, sizeof(V));
  __stcs(reinterpret_cast<V*>(p), v);
}

// One GAE step, written exactly in SB3's evaluation order, every op rounded.
__device__ __forceinline__ void gae_step(float r, float v, float& next_v, float& next_nnt,
                                         float& last, float g, float c, float& adv,
                                         float& ret) {
  float t1 = g * next_v;
  t1 = t1 * next_nnt;
  float d = (r + t1) - v;
  float cc = c * next_nnt;
  last = d + cc * last;
  adv = last;
  ret = last + v;
  next_v = v;
}

template <int VEC, int U>
__global__ void __launch_bounds__(128)
gae_window_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                  const float* __restrict__ start, const float* __restrict__ last_values,
                  const float* __restrict__ dones, float* __restrict__ adv_out,
                  float* __restrict__ ret_out, int64_t T, int64_t N, float g, float c) {
  const int64_t n0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (n0 >= N) return;

  float next_v[VEC], next_nnt[VEC], last[VEC];
  {
    float lv[VEC], dn[VEC];
    ld_stream<VEC>(last_values + n0, lv);
    ld_stream<VEC>(dones + n0, dn);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      next_v[i] = lv[i];
      next_nnt[i] = 1.0f - dn[i];
      last[i] = 0.0f;
    }
  }

  int64_t t = T - 1;
  // full windows
  for (; t >= U - 1; t -= U) {
    float r[U][VEC], v[U][VEC], s[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t off = (t - u) * N + n0;
      ld_stream<VEC>(rew + off, r[u]);
      ld_stream<VEC>(val + off, v[u]);
      ld_stream<VEC>(start + off, s[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t off = (t - u) * N + n0;
      float a[VEC], q[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        gae_step(r[u][i], v[u][i], next_v[i], next_nnt[i], last[i], g, c, a[i], q[i]);
        next_nnt[i] = 1.0f - s[u][i];
      }
      st_stream<VEC>(adv_out + off, a);
      st_stream<VEC>(ret_out + off, q);
    }
  }
  // tail
  for (; t >= 0; --t) {
    const int64_t off = t * N + n0;
    float r[VEC], v[VEC], s[VEC], a[VEC], q[VEC];
    ld_stream<VEC>(rew + off, r);
    ld_stream<VEC>(val + off, v);
    ld_stream<VEC>(start + off, s);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      gae_step(r[i], v[i], next_v[i], next_nnt[i], last[i], g, c, a[i], q[i]);
      next_nnt[i] = 1.0f - s[i];
    }
    st_stream<VEC>(adv_out + off, a);
    st_stream<VEC>(ret_out + off, q);
  }
}

// ---------------------------------------------------------------------------
// Time-parallel variant for small N: one warp per env, each lane owns a
// contiguous chunk of timesteps.  Pass 1 reduces the chunk to the affine map
// A_first = a + b * A_in; a 5-step reverse warp scan composes the maps; pass 2
// replays the chunk with the right A_in.  Agrees with the sequential
// recurrence to fp32 rounding of the composed maps.
__global__ void __launch_bounds__(128)
gae_warpscan_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                    const float* __restrict__ start, const float* __restrict__ last_values,
                    const float* __restrict__ dones, float* __restrict__ adv_out,
                    float* __restrict__ ret_out, int64_t T, int64_t N, float g, float c) {
  const int lane = threadIdx.x & 31;
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;  // warp-uniform
  const int64_t C = (T + 31) / 32;
  const int64_t t_lo = (int64_t)lane * C;
  int64_t t_hi = t_lo + C - 1;  // inclusive
  if (t_hi > T - 1) t_hi = T - 1;
  const bool has = t_lo <= T - 1;

  // values entering this chunk from the future side
  float nv = 0.f, nnt = 0.f;
  if (has) {
    if (t_hi == T - 1) {
      nv = last_values[n];
      nnt = 1.0f - dones[n];
    } else {
      nv = val[(t_hi + 1) * N + n];
      nnt = 1.0f - start[(t_hi + 1) * N + n];
    }
  }
  // pass 1: affine map of the chunk
  float a = 0.f, b = 1.f;
  if (has) {
    float pv = nv, pn = nnt;
    for (int64_t t = t_hi; t >= t_lo; --t) {
      const int64_t off = t * N + n;
      float r = rew[off], v = val[off], s = start[off];
      float t1 = g * pv;
      t1 = t1 * pn;
      float d = (r + t1) - v;
      float cc = c * pn;
      a = d + cc * a;
      b = cc * b;
      pv = v;
      pn = 1.0f - s;
    }
  }
  // reverse inclusive scan over lanes: lane L gets the composition of maps
  // L, L+1, ..., 31 applied to A_in = 0.
  float sa = a, sb = b;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    float oa = __shfl_down_sync(0xffffffffu, sa, d);
    float ob = __shfl_down_sync(0xffffffffu, sb, d);
    if (lane + d < 32) {
      sa = sa + sb * oa;
      sb = sb * ob;
    }
  }
  // A_in of lane L = suffix value of lane L+1 (0 for the last lane)
  float ain = __shfl_down_sync(0xffffffffu, sa, 1);
  if (lane == 31) ain = 0.f;
  // pass 2: replay
  if (has) {
    float pv = nv, pn = nnt, last = ain;
    for (int64_t t = t_hi; t >= t_lo; --t) {
      const int64_t off = t * N + n;
      float r = rew[off], v = val[off], s = start[off];
      float adv, ret;
      gae_step(r, v, pv, pn, last, g, c, adv, ret);
      pn = 1.0f - s;
      adv_out[off] = adv;
      ret_out[off] = ret;
    }
  }
}

// Ragged partner buffer: per env valid prefix [0, count[n]).  Bootstrap rule
// of agents.py:127-129: last_values = values[count-1] (the value stored with
// the last recorded decision), dones = the latched done flag.
__global__ void __launch_bounds__(128)
gae_ragged_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                  const float* __restrict__ start, const int32_t* __restrict__ count,
                  const float* __restrict__ last_done, float* __restrict__ adv_out,
                  float* __restrict__ ret_out, int64_t Tcap, int64_t N, float g, float c) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int64_t cnt = count[n];
  if (cnt <= 0) return;
  if (cnt > Tcap) cnt = Tcap;
  float next_v = val[(cnt - 1) * N + n];
  float next_nnt = 1.0f - last_done[n];
  float last = 0.f;
  // a register window of U timesteps: the 3 * U loads are issued together (they do not depend on
  // the recurrence), then the U steps retire in order — one thread per env has nothing else to
  // hide the DRAM latency with (profiles/kernels_r02.md: 72 cycles of long-scoreboard stall per
  // issued instruction before)
  constexpr int U = 8;
  for (int64_t t = cnt - 1; t >= 0; t -= U) {
    float r[U], v[U], s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t tt = t - u;
      const int64_t off = (tt >= 0 ? tt : 0) * N + n;  // clamped: loaded anyway, used only if tt >= 0
      r[u] = __ldcs(rew + off);
      v[u] = __ldcs(val + off);
      s[u] = __ldcs(start + off);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t tt = t - u;
      if (tt >= 0) {
        float adv, ret;
        gae_step(r[u], v[u], next_v, next_nnt, last, g, c, adv, ret);
        next_nnt = 1.0f - s[u];
        __stcs(adv_out + tt * N + n, adv);
        __stcs(ret_out + tt * N + n, ret);
      }
    }
  }
}

template <int VEC, int U>
int launch_window(const float* rew, const float* val, const float* start, const float* lv,
                  const float* dn, float* adv, float* ret, int64_t T, int64_t N, float g,
                  float c, int block, cudaStream_t st) {
  int64_t threads = (N + VEC - 1) / VEC;
  int grid = pth_ceil_div(threads, block);
  gae_window_kernel<VEC, U><<<grid, block, 0, st>>>(rew, val, start, lv, dn, adv, ret, T, N, g, c);
  return 0;
}

}  // namespace

int pth_gae_tma_launch(pth_ctx* ctx, const float* rew, const float* val, const float* start,
                       const float* lv, const float* dn, float* adv, float* ret, int64_t T,
                       int64_t N, float g, float c, int tune, cudaStream_t st);

extern "C" int pth_gae_f32(pth_ctx* ctx, const float* d_rewards, const float* d_values,
                           const float* d_episode_starts, const float* d_last_values,
                           const float* d_dones, float* d_advantages, float* d_returns,
                           int64_t T, int64_t N, double gamma, double gae_lambda, int variant,
                           void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_rewards && d_values && d_episode_starts && d_last_values && d_dones &&
                    d_advantages && d_returns,
                "NULL device pointer");
  PTH_CHECK_ARG(T >= 0 && N >= 0, "negative size");
  if (T == 0 || N == 0) return PTH_OK;
  PTH_CHECK_ARG(T * N < ((int64_t)1 << 40), "buffer too large");
  cudaStream_t st = (cudaStream_t)stream;
  // SB3 multiplies python doubles first: gamma * gae_lambda, then casts to the
  // array dtype (float32).
  const float g = (float)gamma;
  const float c = (float)(gamma * gae_lambda);

  int v = variant;
  int tune = 0;
  if (variant >= 0x1000) {
    v = (variant >> 12) & 0xf;
    tune = variant & 0xfff;
  }
  if (v == 0) {
    // measured on B200 (profiles/gae_tuning_r01.md): the TMA-staged kernel wins
    // once every SM has a few hundred envs; below that the register-window
    // kernel has lower fixed cost; tiny N is latency bound -> time-parallel scan.
    if (N < 512 && T >= 64)
      v = 3;
    else if (N % 4 == 0 && N >= (int64_t)ctx->sm_count * 128 &&
             (((uintptr_t)d_rewards | (uintptr_t)d_values | (uintptr_t)d_episode_starts) % 16) == 0)
      v = 2;
    else
      v = 1;
  }
  const bool al4 = (N % 4 == 0) && (((uintptr_t)d_rewards | (uintptr_t)d_values |
                                      (uintptr_t)d_episode_starts | (uintptr_t)d_last_values |
                                      (uintptr_t)d_dones | (uintptr_t)d_advantages |
                                      (uintptr_t)d_returns) % 16 == 0);
  const bool al2 = (N % 2 == 0) && (((uintptr_t)d_rewards | (uintptr_t)d_values |
                                      (uintptr_t)d_episode_starts | (uintptr_t)d_last_values |
                                      (uintptr_t)d_dones | (uintptr_t)d_advantages |
                                      (uintptr_t)d_returns) % 8 == 0);
  if (v == 1) {
    // tune: bits 8..11 vec (1,2,4), bits 4..7 window code (1:4 2:8 3:16), bits 0..3 block/32
    int vec = (tune >> 8) & 0xf, ucode = (tune >> 4) & 0xf, blk = (tune & 0xf) * 32;
    int U = ucode == 0 ? 0 : (ucode == 1 ? 4 : (ucode == 2 ? 8 : (ucode == 3 ? 16 : -1)));
    if (blk > 128 || U < 0) {
      pth_set_error("pth_gae_f32: bad tuning code 0x%x", variant);
      return PTH_EINVAL;
    }
    if (vec == 0) {
      // auto: keep >= ~2 warps per SM worth of threads before widening loads
      const int64_t sm = ctx->sm_count;
      if (al4 && N >= sm * 128 * 4)
        vec = 4;
      else if (al2 && N >= sm * 64 * 2)
        vec = 2;
      else
        vec = 1;
    }
    if (vec == 4 && !al4) vec = al2 ? 2 : 1;
    if (vec == 2 && !al2) vec = 1;
    if (U == 0) U = 8;
    if (blk == 0) blk = 64;
#define PTH_GAE_CASE(VV, UU)                                                              \
  if (vec == VV && U == UU) {                                                             \
    launch_window<VV, UU>(d_rewards, d_values, d_episode_starts, d_last_values, d_dones,  \
                          d_advantages, d_returns, T, N, g, c, blk, st);                  \
  } else
    PTH_GAE_CASE(1, 4)
    PTH_GAE_CASE(1, 8)
    PTH_GAE_CASE(1, 16)
    PTH_GAE_CASE(2, 4)
    PTH_GAE_CASE(2, 8)
    PTH_GAE_CASE(2, 16)
    PTH_GAE_CASE(4, 4)
    PTH_GAE_CASE(4, 8)
    PTH_GAE_CASE(4, 16) {
      pth_set_error("pth_gae_f32: bad tuning code 0x%x", variant);
      return PTH_EINVAL;
    }
#undef PTH_GAE_CASE
  } else if (v == 2) {
    int rc = pth_gae_tma_launch(ctx, d_rewards, d_values, d_episode_starts, d_last_values,
                                d_dones, d_advantages, d_returns, T, N, g, c, tune, st);
    if (rc != PTH_OK) return rc;
  } else if (v == 3) {
    int64_t threads = N * 32;
    gae_warpscan_kernel<<<pth_ceil_div(threads, 128), 128, 0, st>>>(
        d_rewards, d_values, d_episode_starts, d_last_values, d_dones, d_advantages, d_returns,
        T, N, g, c);
  } else {
    pth_set_error("pth_gae_f32: unknown variant %d", variant);
    return PTH_EINVAL;
  }
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_gae_ragged_f32(pth_ctx* ctx, const float* d_rewards, const float* d_values,
                                  const float* d_episode_starts, const int32_t* d_count,
                                  const float* d_last_done, float* d_advantages,
                                  float* d_returns, int64_t Tcap, int64_t N, double gamma,
                                  double gae_lambda, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_rewards && d_values && d_episode_starts && d_count && d_last_done &&
                    d_advantages && d_returns,
                "NULL device pointer");
  PTH_CHECK_ARG(Tcap >= 0 && N >= 0, "negative size");
  if (Tcap == 0 || N == 0) return PTH_OK;
  const float g = (float)gamma;
  const float c = (float)(gamma * gae_lambda);
  gae_ragged_kernel<<<pth_ceil_div(N, 32), 32, 0, (cudaStream_t)stream>>>(
      d_rewards, d_values, d_episode_starts, d_count, d_last_done, d_advantages, d_returns, Tcap,
      N, g, c);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}
