// pth_api.cu — context, error reporting and space helpers of the C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "pth_common.cuh"

static thread_local char g_err[512] = "";

// FP32 FFMA peak probe (pth_debug_ffma_peak): 8 independent chains per thread hide the 4-cycle FFMA
// latency with 32 warps per SM; nothing but FFMA in the loop body.
__global__ void __launch_bounds__(1024) ffma_peak_kernel(float* sink, int iters) {
  const float t = (float)threadIdx.x * 1e-6f;
  float a0 = t, a1 = t + 1.f, a2 = t + 2.f, a3 = t + 3.f, a4 = t + 4.f, a5 = t + 5.f, a6 = t + 6.f, a7 = t + 7.f;
  const float m = 0.999f + t, c = 1e-3f;
#pragma unroll 1
  for (int i = 0; i < iters; i += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  }
  sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

void pth_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

int pth_version(void) { return PTH_VERSION; }

const char* pth_last_error(void) { return g_err; }

int pth_ctx_create(int device, pth_ctx** out) {
  PTH_CHECK_ARG(out != nullptr, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    pth_set_error("pth_ctx_create: no CUDA device (%s); this library has no CPU path",
                  e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    return PTH_ENODEV;
  }
  PTH_CHECK_ARG(device >= 0 && device < count, "device index out of range");
  cudaDeviceProp prop;
  PTH_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    pth_set_error("pth_ctx_create: device %d is sm_%d%d; kernels are built for sm_100a only",
                  device, prop.major, prop.minor);
    return PTH_ENODEV;
  }
  pth_ctx* c = (pth_ctx*)calloc(1, sizeof(pth_ctx));
  PTH_CHECK_ARG(c != nullptr, "out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->coop_launch = prop.cooperativeLaunch;
  *out = c;
  return PTH_OK;
}

int pth_ctx_destroy(pth_ctx* ctx) {
  if (ctx) {
    pth_comm_destroy(ctx);
    free(ctx);
  }
  return PTH_OK;
}

int pth_ctx_sm_count(const pth_ctx* ctx) { return ctx ? ctx->sm_count : PTH_EINVAL; }

int pth_sync_debug(pth_ctx* ctx, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  PTH_CUDA(cudaGetLastError());
  return PTH_OK;
}

int pth_debug_ffma_peak(pth_ctx* ctx, float* d_sink, int32_t ctas, int32_t iters, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && d_sink != nullptr && ctas > 0 && iters > 0 && iters % 8 == 0,
                "NULL ctx / sink, or iters not a positive multiple of 8");
  ffma_peak_kernel<<<ctas, 1024, 0, (cudaStream_t)stream>>>(d_sink, iters);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

static int space_ok(const pth_space* sp) {
  if (!sp) return 0;
  if (sp->obs_kind != PTH_OBS_ONEHOT && sp->obs_kind != PTH_OBS_BOX) return 0;
  if (sp->obs_len < 1 || sp->obs_len > PTH_MAX_OBS_SLOTS) return 0;
  if (sp->n_heads < 1 || sp->n_heads > PTH_MAX_HEADS) return 0;
  for (int h = 0; h < sp->n_heads; ++h)
    if (sp->head_n[h] < 1 || sp->head_n[h] > 32) return 0;
  if (sp->obs_kind == PTH_OBS_ONEHOT)
    for (int s = 0; s < sp->obs_len; ++s)
      if (sp->obs_nvec[s] < 1 || sp->obs_nvec[s] > 255) return 0;
  return 1;
}

int pth_space_feature_dim(const pth_space* sp) {
  if (!space_ok(sp)) return PTH_EINVAL;
  if (sp->obs_kind == PTH_OBS_BOX) return sp->obs_len;
  int f = 0;
  for (int s = 0; s < sp->obs_len; ++s) f += sp->obs_nvec[s];
  return f;
}

int pth_space_logit_dim(const pth_space* sp) {
  if (!space_ok(sp)) return PTH_EINVAL;
  int l = 0;
  for (int h = 0; h < sp->n_heads; ++h) l += sp->head_n[h];
  return l;
}

int64_t pth_policy_param_count(const pth_space* sp) {
  if (!space_ok(sp)) return PTH_EINVAL;
  int64_t F = pth_space_feature_dim(sp), L = pth_space_logit_dim(sp), H = PTH_HIDDEN;
  return 2 * (H * F + H + H * H + H) + L * H + L + H + 1;
}

int64_t pth_adap_param_count(const pth_space* sp, int32_t context_size) {
  if (!space_ok(sp) || context_size < 0 || context_size > 8) return PTH_EINVAL;
  return pth_policy_param_count(sp) + 2 * (int64_t)PTH_HIDDEN * context_size;
}

}  // extern "C"
