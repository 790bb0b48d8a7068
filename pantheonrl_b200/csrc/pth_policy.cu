// pth_policy.cu — a1 (+ the value/log-prob part of a2): batched policy forward,
// categorical sampling, log-prob, entropy, value.
//
// Replaces util.action_from_policy (pantheonrl/common/util.py:63-81) ->
// ActorCriticPolicy.forward and, with d_action_in, evaluate_actions.
#include "pth_mlp.cuh"

using namespace pthmlp;

namespace {

struct FwdSmem {
  SmemPolicy pol;
  float A[HID * LDA];
  float Bf[HID * LDA];
  float Lg[MAXL * LDA];
  uint8_t obs[BT * 32];
  // Box observations only: X[F][LDA] follows (F <= 64); wide one-hot rows (more than 32 slots):
  // the [BT][96] byte tile follows instead of `obs`
};

struct FwdParams {
  SpaceDev sp;
  Layout lo;
  const float* params;
  const void* obs;
  int64_t obs_stride;
  int64_t B;
  uint64_t seed;
  uint32_t rng_stream, tick, slot;
  int64_t idx0;
  const uint8_t* action_in;
  uint8_t* action;
  float* value;
  float* logp;
  float* entropy;
  float* logits;
  const float* race;
  // AdapPolicy (pantheonrl/algos/adap/policies.py:86-131): C context inputs behind the features
  int C;
  const float* ctx;
  int64_t ctx_stride;
  int cx_off;  // byte offset of the [8][LDA] context tile in dynamic shared memory
  // ModularPolicy (pantheonrl/algos/modular/policies.py:273-290, 364-383): Pn partner modules behind the
  // main network, module p0 composes this forward: logits = main + partner, value = main + partner
  int Pn, p0;
  int x_off;  // byte offset of one more [64][LDA] tile (the partner branches' second layer)
  // AdapPolicyMult (adap/policies.py:134-283): scaling layers' offsets (tower 0 = policy), x_off tile = a scaling
  // sub-layer's output, w_off = a [64][LDW] weight slot + 64 biases
  int mult, mw[2], mb[2], w_off;
};

// a [rows][64] matrix / a vector from global memory into a shared-memory slot (row stride LDW); the partner
// blocks follow the main network without padding, so they need not be 16-byte aligned: scalar loads
__device__ __forceinline__ void stage_rows(float* dst, const float* src, int rows, int tid) {
  for (int i = tid; i < rows * HID; i += NT) dst[(i >> 6) * LDW + (i & 63)] = __ldg(src + i);
}
__device__ __forceinline__ void stage_v(float* dst, const float* src, int n, int tid) {
  for (int i = tid; i < n; i += NT) dst[i] = __ldg(src + i);
}

template <int OW>
__global__ void __launch_bounds__(NT) policy_forward_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  float* Xs = reinterpret_cast<float*>(smem_raw + sizeof(FwdSmem));
  uint8_t* obs_s = OW == 32 ? sm.obs : smem_raw + sizeof(FwdSmem);
  float* Cx = reinterpret_cast<float*>(smem_raw + p.cx_off);
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * BT;
  const int64_t b = b0 + tid;
  const bool lane = tid < BT;          // threads [0, BT) own one sample each
  const bool live = lane && b < p.B;   // threads [BT, NT) only help in the tiled layers

  load_policy(sm.pol, p.params, p.lo, p.sp.L, tid, NT);
  if (lane) {
    if (p.sp.obs_kind == PTH_OBS_ONEHOT) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.obs);
      for (int s = 0; s < OW; ++s) {
        uint8_t v = 0;
        if (live && s < p.sp.obs_len) v = src[b * p.obs_stride + s];
        obs_s[tid * OW + s] = v;
      }
    } else {
      const float* src = reinterpret_cast<const float*>(p.obs);
      for (int k = 0; k < p.sp.F; ++k) Xs[k * LDA + tid] = live ? src[b * p.obs_stride + k] : 0.f;
    }
    for (int c = 0; c < p.C; ++c) Cx[c * LDA + tid] = live ? p.ctx[b * p.ctx_stride + c] : 0.f;
  }
  __syncthreads();
  // first layer of one tower into sm.A; AdapPolicy: the context inputs continue the chains, then tanh
  const bool ctx_cols = p.C > 0 && !p.mult;  // AdapPolicy: context columns; AdapPolicyMult: features only
  auto first_layer = [&](int w_off, const float* bias) {
    if (p.sp.obs_kind == PTH_OBS_ONEHOT)
      first_layer_onehot<false, NT, BT, 6, OW>(p.sp, obs_s, p.params + w_off, bias, sm.A, tid, !ctx_cols);
    else if (!ctx_cols)
      first_layer_box<false>(p.sp.F, Xs, p.params + w_off, bias, sm.A, tid);
    else
      first_layer_box<false, NT, BT, false>(p.sp.F, Xs, p.params + w_off, bias, sm.A, tid);
    if (ctx_cols) {
      __syncthreads();
      context_columns_tanh<false>(p.C, Cx, p.params + w_off + p.sp.F * HID, sm.A, tid >> 5, NT / 32, tid & 31);
    }
  };

  float v = 0.f;
  if (p.mult) {
    // AdapPolicyMult tower: x = first layer (sm.A); y = x + sum_c s_c ctx_c (sm.Bf), s_c = tanh(Ws_c x + bs_c)
    // (rows j C + c of the scaling layer, staged sub-layer by sub-layer); h2 = tanh(W1 y + b1) (sm.A again)
    float* X = reinterpret_cast<float*>(smem_raw + p.x_off);
    float* Wslot = reinterpret_cast<float*>(smem_raw + p.w_off);
    float* Wbias = Wslot + HID * LDW;
    for (int t = 0; t < 2; ++t) {
      first_layer(t ? p.lo.w_vf0 : p.lo.w_pi0, t ? sm.pol.b_vf0 : sm.pol.b_pi0);
      __syncthreads();
      for (int i = tid; i < HID * BT; i += NT) sm.Bf[(i >> 7) * LDA + (i & (BT - 1))] = sm.A[(i >> 7) * LDA + (i & (BT - 1))];
      for (int c = 0; c < p.C; ++c) {
        for (int i = tid; i < HID * HID; i += NT)
          Wslot[(i >> 6) * LDW + (i & 63)] = __ldg(p.params + p.mw[t] + ((size_t)(i >> 6) * p.C + c) * HID + (i & 63));
        for (int i = tid; i < HID; i += NT) Wbias[i] = __ldg(p.params + p.mb[t] + (size_t)i * p.C + c);
        __syncthreads();
        dense64<true>(sm.A, Wslot, Wbias, X, tid);
        __syncthreads();
        for (int i = tid; i < HID * BT; i += NT) {
          const int k = i >> 7, b = i & (BT - 1);
          sm.Bf[k * LDA + b] = fmaf(X[k * LDA + b], Cx[c * LDA + b], sm.Bf[k * LDA + b]);
        }
        __syncthreads();
      }
      dense64<true>(sm.Bf, t ? sm.pol.w_vf1 : sm.pol.w_pi1, t ? sm.pol.b_vf1 : sm.pol.b_pi1, sm.A, tid);
      __syncthreads();
      if (lane) {
        if (t == 0)
          action_head(sm.A, sm.pol, p.sp.L, sm.Lg, tid);
        else
          v = value_head(sm.A, sm.pol, tid);
      }
      __syncthreads();
    }
    if (!lane) return;
  } else {
  // ---- policy tower
  first_layer(p.lo.w_pi0, sm.pol.b_pi0);
  __syncthreads();
  dense64<true>(sm.A, sm.pol.w_pi1, sm.pol.b_pi1, sm.Bf, tid);
  __syncthreads();
  if (lane) action_head(sm.Bf, sm.pol, p.sp.L, sm.Lg, tid);
  // ---- value tower (A is free again)
  if (p.Pn == 0) {
    first_layer(p.lo.w_vf0, sm.pol.b_vf0);
    __syncthreads();
    dense64<true>(sm.A, sm.pol.w_vf1, sm.pol.b_vf1, sm.Bf, tid);
    __syncthreads();
    if (!lane) return;
    v = value_head(sm.Bf, sm.pol, tid);
  } else {
    // ModularPolicy: sm.Bf keeps latent_pi for the partner module; the extra tile X takes the second layers
    float* X = reinterpret_cast<float*>(smem_raw + p.x_off);
    const int L = p.sp.L;
    const int blk = 4 * (HID * HID + HID) + L * HID + L + HID + 1;  // oracle/pth_oracle_modular.inc: mod_block
    const float* pb = p.params + p.lo.total + (size_t)p.p0 * blk;
    const float* pw_pi0 = pb, *pb_pi0 = pw_pi0 + HID * HID, *pw_pi1 = pb_pi0 + HID, *pb_pi1 = pw_pi1 + HID * HID;
    const float* pw_vf0 = pb_pi1 + HID, *pb_vf0 = pw_vf0 + HID * HID, *pw_vf1 = pb_vf0 + HID, *pb_vf1 = pw_vf1 + HID * HID;
    const float* pw_act = pb_vf1 + HID, *pb_act = pw_act + L * HID, *pw_val = pb_act + L, *pb_val = pw_val + HID;
    first_layer(p.lo.w_vf0, sm.pol.b_vf0);
    __syncthreads();
    dense64<true>(sm.A, sm.pol.w_vf1, sm.pol.b_vf1, X, tid);
    __syncthreads();
    if (lane) v = value_head(X, sm.pol, tid);
    __syncthreads();
    // partner policy branch: q1 = tanh(W0 latent_pi + b), q2 = tanh(W1 q1 + b), logits += Wact q2 + bact
    stage_rows(sm.pol.w_pi1, pw_pi0, HID, tid);
    stage_rows(sm.pol.w_vf1, pw_pi1, HID, tid);
    stage_rows(sm.pol.w_act, pw_act, L, tid);
    stage_v(sm.pol.b_pi1, pb_pi0, HID, tid);
    stage_v(sm.pol.b_vf1, pb_pi1, HID, tid);
    stage_v(sm.pol.b_act, pb_act, L, tid);
    __syncthreads();
    dense64<true>(sm.Bf, sm.pol.w_pi1, sm.pol.b_pi1, sm.A, tid);
    __syncthreads();
    dense64<true>(sm.A, sm.pol.w_vf1, sm.pol.b_vf1, X, tid);
    __syncthreads();
    if (lane) {
      float h[HID];
      load_column(X, tid, h);
      for (int l = 0; l < L; ++l) sm.Lg[l * LDA + tid] = sm.Lg[l * LDA + tid] + dot64(h, sm.pol.w_act + l * LDW, sm.pol.b_act[l]);
    }
    __syncthreads();
    // partner value branch (it reads latent_pi too)
    stage_rows(sm.pol.w_pi1, pw_vf0, HID, tid);
    stage_rows(sm.pol.w_vf1, pw_vf1, HID, tid);
    stage_v(sm.pol.b_pi1, pb_vf0, HID, tid);
    stage_v(sm.pol.b_vf1, pb_vf1, HID, tid);
    stage_v(sm.pol.w_val, pw_val, HID, tid);
    if (tid == 0) sm.pol.b_val = __ldg(pb_val);
    __syncthreads();
    dense64<true>(sm.Bf, sm.pol.w_pi1, sm.pol.b_pi1, sm.A, tid);
    __syncthreads();
    dense64<true>(sm.A, sm.pol.w_vf1, sm.pol.b_vf1, X, tid);
    __syncthreads();
    if (!lane) return;
    v = v + value_head(X, sm.pol, tid);
  }
  }

  // ---- distribution
  bool sample = p.action_in == nullptr;
  pth_u4 rnd = {0, 0, 0, 0};
  uint32_t ain = 0;
  if (sample && p.race != nullptr) {
    // exponential race per head (torch.multinomial's single-sample path): argmax_i p_i / q_i,
    // first maximum wins; p_i as dist_eval computes it (exp(z_i - max) / S)
    if (live) {
      const float* q = p.race + b * p.sp.L;
      int off = 0;
      for (int h = 0; h < p.sp.n_heads; ++h) {
        const int n = p.sp.head_n[h];
        float m = sm.Lg[off * LDA + tid];
        for (int i = 1; i < n; ++i) {
          const float z = sm.Lg[(off + i) * LDA + tid];
          m = z > m ? z : m;
        }
        float S = 0.f;
        for (int i = 0; i < n; ++i) S = S + pth_expf(sm.Lg[(off + i) * LDA + tid] - m);
        float best = -1.0f;
        int a = 0;
        for (int i = 0; i < n; ++i) {
          const float r = (pth_expf(sm.Lg[(off + i) * LDA + tid] - m) / S) / q[off + i];
          if (r > best) {
            best = r;
            a = i;
          }
        }
        ain |= ((uint32_t)a & 0xffu) << (8 * h);
        off += n;
      }
    }
    sample = false;
  } else if (sample) {
    rnd = pth_philox(p.seed, p.rng_stream, (uint64_t)(p.idx0 + b), p.tick, p.slot);
  } else if (live) {
    ain = *reinterpret_cast<const uint32_t*>(p.action_in + 4 * b);
  }
  DistOut d = dist_eval(p.sp, sm.Lg, tid, sample, rnd, ain);
  if (!live) return;
  if (p.action) *reinterpret_cast<uint32_t*>(p.action + 4 * b) = d.action;
  if (p.value) p.value[b] = v;
  if (p.logp) p.logp[b] = d.logp;
  if (p.entropy) p.entropy[b] = d.entropy;
  if (p.logits)
    for (int l = 0; l < p.sp.L; ++l) p.logits[b * p.sp.L + l] = sm.Lg[l * LDA + tid];
}

__global__ void math_kernel(int which, const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  y[i] = which == 0 ? pth_expf(v) : (which == 1 ? pth_logf(v) : pth_tanhf(v));
}

}  // namespace

extern "C" int pth_debug_math(pth_ctx* ctx, int which, const float* d_x, float* d_y, int64_t n, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && d_x && d_y && n >= 0 && which >= 0 && which <= 2, "bad argument");
  if (n == 0) return PTH_OK;
  math_kernel<<<pth_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(which, d_x, d_y, n);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_policy_forward(pth_ctx* ctx, const pth_forward_args* a, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && a != nullptr, "NULL ctx/args");
  PTH_CHECK_ARG(a->space != nullptr && a->d_params != nullptr && a->d_obs != nullptr,
                "NULL space/params/obs");
  PTH_CHECK_ARG(a->B >= 0, "negative batch");
  if (a->B == 0) return PTH_OK;
  FwdParams p;
  if (fill_space(a->space, &p.sp) != 0) {
    pth_set_error("pth_policy_forward: unsupported space");
    return PTH_ENOSUP;
  }
  if (p.sp.obs_kind == PTH_OBS_BOX && p.sp.F > HID) {
    pth_set_error("pth_policy_forward: Box observations wider than 64 are not supported");
    return PTH_ENOSUP;
  }
  PTH_CHECK_ARG(a->obs_stride >= p.sp.obs_len, "obs_stride smaller than obs_len");
  PTH_CHECK_ARG(((uintptr_t)a->d_params % 16) == 0, "params must be 16-byte aligned");
  PTH_CHECK_ARG(a->context_size >= 0 && a->context_size <= 8 && (a->context_size == 0 || a->d_context != nullptr),
                "context_size 0..8 with d_context");
  PTH_CHECK_ARG(a->context_stride == 0 || a->context_stride >= a->context_size, "context_stride smaller than context_size");
  p.C = a->context_size;
  p.ctx = a->d_context;
  p.ctx_stride = a->context_stride;
  p.lo = make_layout(p.sp.F + p.C, p.sp.L);
  p.mult = a->adap_mult != 0;
  p.mw[0] = p.mw[1] = p.mb[0] = p.mb[1] = 0;
  if (p.mult) {  // per tower: first layer (features only) | scaling 64 -> 64 C | second layer; then the heads
    PTH_CHECK_ARG(p.C > 0 && a->num_partners == 0, "AdapPolicyMult needs context_size > 0 (and no partner modules)");
    int o = 0;
    const int F = p.sp.F, L = p.sp.L, C = p.C;
    p.lo.w_pi0 = o; o += HID * F;
    p.lo.b_pi0 = o; o += HID;
    p.mw[0] = o; o += HID * C * HID;
    p.mb[0] = o; o += HID * C;
    p.lo.w_pi1 = o; o += HID * HID;
    p.lo.b_pi1 = o; o += HID;
    p.lo.w_vf0 = o; o += HID * F;
    p.lo.b_vf0 = o; o += HID;
    p.mw[1] = o; o += HID * C * HID;
    p.mb[1] = o; o += HID * C;
    p.lo.w_vf1 = o; o += HID * HID;
    p.lo.b_vf1 = o; o += HID;
    p.lo.w_act = o; o += L * HID;
    p.lo.b_act = o; o += L;
    p.lo.w_val = o; o += HID;
    p.lo.b_val = o; o += 1;
    p.lo.total = o;
  }
  p.params = a->d_params;
  p.obs = a->d_obs;
  p.obs_stride = a->obs_stride;
  p.B = a->B;
  p.seed = a->seed;
  p.rng_stream = a->rng_stream;
  p.tick = a->tick;
  p.slot = a->slot;
  p.idx0 = a->idx0;
  p.action_in = a->d_action_in;
  p.action = a->d_action;
  p.value = a->d_value;
  p.logp = a->d_logp;
  p.entropy = a->d_entropy;
  p.logits = a->d_logits;
  p.race = a->d_race;
  const bool wide = p.sp.obs_kind == PTH_OBS_ONEHOT && p.sp.obs_len > 32;
  size_t smem = sizeof(FwdSmem) + (p.sp.obs_kind == PTH_OBS_BOX ? sizeof(float) * HID * LDA : 0) + (wide ? BT * 96 : 0);
  p.cx_off = (int)smem;
  smem += p.C > 0 ? sizeof(float) * 8 * LDA : 0;
  PTH_CHECK_ARG(a->num_partners >= 0 && a->num_partners <= 8 && (a->num_partners == 0 ||
                (a->partner_idx >= 0 && a->partner_idx < a->num_partners && a->context_size == 0)),
                "ModularPolicy: 1..8 partners, partner_idx among them, no context inputs");
  p.Pn = a->num_partners;
  p.p0 = a->partner_idx;
  p.x_off = (int)smem;
  smem += (p.Pn > 0 || p.mult) ? sizeof(float) * HID * LDA : 0;
  p.w_off = (int)smem;
  smem += p.mult ? sizeof(float) * (HID * LDW + HID) : 0;
  if (wide) {
    PTH_CUDA(cudaFuncSetAttribute(policy_forward_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    policy_forward_kernel<96><<<pth_ceil_div(a->B, BT), NT, smem, (cudaStream_t)stream>>>(p);
  } else {
    PTH_CUDA(cudaFuncSetAttribute(policy_forward_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    policy_forward_kernel<32><<<pth_ceil_div(a->B, BT), NT, smem, (cudaStream_t)stream>>>(p);
  }
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}
