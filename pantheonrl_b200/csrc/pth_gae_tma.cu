// pth_gae_tma.cu — GAE with rollout rows staged into shared memory by the TMA
// engine (cp.async.bulk, mbarrier complete_tx), sm_100a.
//
// Each CTA owns a contiguous slab of W envs for the whole horizon.  A producer
// warp walks t = T-1 .. 0 in blocks of TT rows and issues one bulk copy per
// (array, row) into a ring of STAGES shared-memory stages; 256 consumer threads
// (one env each) retire a stage with the exact sequential recurrence and store
// advantages / returns straight to global memory (coalesced, streaming).  The
// loads in flight per SM are STAGES * TT * 3 * W * 4 bytes, independent of how
// many threads are resident — this is what the register-window kernel cannot
// do when N gives it only a few warps per SM.
#include "pth_common.cuh"

namespace {

constexpr int kConsumers = 256;
constexpr int kThreads = kConsumers + 32;

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pth_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pth_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pth_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(pth_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(pth_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(pth_smem_u32(bar))
      : "memory");
}

template <int TT, int STAGES>
__global__ void __launch_bounds__(kThreads)
gae_tma_kernel(const float* __restrict__ rew, const float* __restrict__ val,
               const float* __restrict__ start, const float* __restrict__ last_values,
               const float* __restrict__ dones, float* __restrict__ adv_out,
               float* __restrict__ ret_out, int64_t T, int64_t N, int W, float g, float c) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // stage layout: [3 arrays][TT rows][W floats]
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  const int stage_floats = 3 * TT * W;
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + (size_t)STAGES * stage_floats);
  uint64_t* empty = full + STAGES;

  const int64_t n0 = (int64_t)blockIdx.x * W;
  if (n0 >= N) return;
  const int w = (int)((N - n0 < W) ? (N - n0) : W);  // multiple of 4
  const int tid = threadIdx.x;
  const int n_consumer_warps = (w + 31) / 32;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], n_consumer_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t n_blocks = (T + TT - 1) / TT;

  if (tid >= kConsumers) {
    // ---------------- producer warp
    const int lane = tid - kConsumers;
    const uint32_t row_bytes = (uint32_t)w * 4u;
    for (int64_t k = 0; k < n_blocks; ++k) {
      const int s = (int)(k % STAGES);
      const uint32_t ph = (uint32_t)((k / STAGES) & 1);
      const int64_t t_hi = T - 1 - k * TT;
      const int rows = (int)((t_hi + 1 < TT) ? (t_hi + 1) : TT);
      if (lane == 0) {
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], (uint32_t)rows * 3u * row_bytes);
      }
      __syncwarp();
      float* sb = stage_base + (size_t)s * stage_floats;
      for (int i = lane; i < rows * 3; i += 32) {
        const int a = i / rows, u = i - a * rows;
        const float* src = (a == 0 ? rew : (a == 1 ? val : start)) + (t_hi - u) * N + n0;
        bulk_g2s(sb + ((size_t)a * TT + u) * W, src, row_bytes, &full[s]);
      }
    }
  } else if ((tid & ~31) < w) {
    // ---------------- consumers: one env each (whole warps take part in the
    // barrier protocol; lanes past the slab edge only skip memory traffic)
    const bool active = tid < w;
    const int64_t n = n0 + (active ? tid : 0);
    float next_v = __ldcs(last_values + n);
    float next_nnt = 1.0f - __ldcs(dones + n);
    float last = 0.f;
    const int lane = tid & 31;
    for (int64_t k = 0; k < n_blocks; ++k) {
      const int s = (int)(k % STAGES);
      const uint32_t ph = (uint32_t)((k / STAGES) & 1);
      const int64_t t_hi = T - 1 - k * TT;
      const int rows = (int)((t_hi + 1 < TT) ? (t_hi + 1) : TT);
      mbar_wait(&full[s], ph);
      const float* sb = stage_base + (size_t)s * stage_floats;
      float r[TT], v[TT], st[TT];
      const int col = active ? tid : 0;
#pragma unroll
      for (int u = 0; u < TT; ++u) {
        if (u < rows) {
          r[u] = sb[(0 * TT + u) * W + col];
          v[u] = sb[(1 * TT + u) * W + col];
          st[u] = sb[(2 * TT + u) * W + col];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);  // stage is in registers now
#pragma unroll
      for (int u = 0; u < TT; ++u) {
        if (u < rows) {
          float t1 = g * next_v;
          t1 = t1 * next_nnt;
          float d = (r[u] + t1) - v[u];
          float cc = c * next_nnt;
          last = d + cc * last;
          const int64_t off = (t_hi - u) * N + n;
          if (active) {
            __stcs(adv_out + off, last);
            __stcs(ret_out + off, last + v[u]);
          }
          next_v = v[u];
          next_nnt = 1.0f - st[u];
        }
      }
    }
  }
}

template <int TT, int STAGES>
int launch(pth_ctx* ctx, const float* rew, const float* val, const float* start,
           const float* lv, const float* dn, float* adv, float* ret, int64_t T, int64_t N,
           int W, float g, float c, cudaStream_t st) {
  size_t smem = (size_t)STAGES * 3 * TT * W * sizeof(float) + 2 * STAGES * sizeof(uint64_t);
  if ((int)smem > ctx->max_smem_optin) {
    pth_set_error("pth_gae_f32(tma): stage ring %zu B exceeds shared memory", smem);
    return PTH_EINVAL;
  }
  auto kern = gae_tma_kernel<TT, STAGES>;
  PTH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = pth_ceil_div(N, W);
  kern<<<grid, kThreads, smem, st>>>(rew, val, start, lv, dn, adv, ret, T, N, W, g, c);
  return PTH_OK;
}

}  // namespace

// tune: bits 8..11 = CTAs per SM target (0 -> 3), bits 4..7 = TT code (0 -> 4;
// 1:4 2:8 3:16 4:2), bits 0..3 = stages (0 -> 8).  Defaults = best measured on
// B200 at (T, N) = (2048, 65536): 6155 GB/s.
int pth_gae_tma_launch(pth_ctx* ctx, const float* rew, const float* val, const float* start,
                       const float* lv, const float* dn, float* adv, float* ret, int64_t T,
                       int64_t N, float g, float c, int tune, cudaStream_t st) {
  if (N % 4 != 0 || (((uintptr_t)rew | (uintptr_t)val | (uintptr_t)start) % 16) != 0) {
    pth_set_error("pth_gae_f32(tma): needs N %% 4 == 0 and 16-byte aligned arrays");
    return PTH_ENOSUP;
  }
  int per_sm = (tune >> 8) & 0xf, ttc = (tune >> 4) & 0xf, stages = tune & 0xf;
  if (per_sm == 0) per_sm = 3;
  if (ttc == 0) ttc = 1;
  if (stages == 0) stages = 8;
  // slab width: spread N over (per_sm * SMs) CTAs in as few equal waves as
  // possible, W a multiple of 4 and <= 256 consumers.
  const int64_t G = (int64_t)per_sm * ctx->sm_count;
  int64_t waves = 1;
  int64_t W = 0;
  for (;; ++waves) {
    W = ((N + G * waves - 1) / (G * waves) + 3) / 4 * 4;
    if (W <= kConsumers) break;
  }
  if (W < 4) W = 4;
#define PTH_TMA_CASE(TTV, SV)                                                         \
  if (ttc == (TTV == 4 ? 1 : (TTV == 8 ? 2 : (TTV == 16 ? 3 : 4))) && stages == SV)  \
    return launch<TTV, SV>(ctx, rew, val, start, lv, dn, adv, ret, T, N, (int)W, g, c, st);
  PTH_TMA_CASE(2, 8)
  PTH_TMA_CASE(2, 12)
  PTH_TMA_CASE(4, 2)
  PTH_TMA_CASE(4, 4)
  PTH_TMA_CASE(4, 6)
  PTH_TMA_CASE(4, 8)
  PTH_TMA_CASE(4, 12)
  PTH_TMA_CASE(8, 2)
  PTH_TMA_CASE(8, 3)
  PTH_TMA_CASE(8, 4)
  PTH_TMA_CASE(8, 6)
  PTH_TMA_CASE(16, 2)
  PTH_TMA_CASE(16, 3)
  PTH_TMA_CASE(16, 4)
#undef PTH_TMA_CASE
  pth_set_error("pth_gae_f32(tma): bad tuning code 0x%x", tune);
  return PTH_EINVAL;
}
