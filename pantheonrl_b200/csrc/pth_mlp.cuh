// pth_mlp.cuh — CTA-cooperative building blocks of the SB3 MlpPolicy
// (two separate 64-64 tanh towers + categorical heads) for tiles of BT = 128
// samples held in shared memory.  Used by the forward kernel, the rollout
// megakernel and the PPO update kernel, so all three produce identical bits.
//
// Reference semantics: stable-baselines3 1.7.0 ActorCriticPolicy.forward /
// evaluate_actions as called from pantheonrl/common/util.py:63-81 (restated
// in-tree at pantheonrl/algos/modular/policies.py:273-290, 364-383).
//
// Numeric contract: every linear output is acc = bias; for k ascending:
// acc = fma(x_k, w_jk, acc); one-hot first layers add their selected rows in descending slot order.  Thread tiles only change WHO computes an output,
// never the order of its additions.
#pragma once
#include "pth_common.cuh"

namespace pthmlp {

constexpr int BT = 128;    // samples per tile
constexpr int NT = 256;    // threads cooperating on a tile (2 warps per SM sub-partition)
constexpr int HID = 64;
constexpr int LDA = 132;   // activation row stride (floats): [feature][sample]
constexpr int LDW = 68;    // hidden weight row stride (floats): [out][in]
constexpr int MAXL = 32;   // max total logits

// Device copy of pth_space with slot offsets precomputed.
struct SpaceDev {
  int obs_kind, obs_len, n_heads, F, L;
  int head_n[PTH_MAX_HEADS];
  int16_t slot_off[PTH_MAX_OBS_SLOTS];
};

// Offsets (floats) into the flat parameter vector.  NOTE: pi0.w / vf0.w are
// stored INPUT-MAJOR [F][64] (transpose of torch's nn.Linear.weight) so that a
// one-hot row gather reads 256 contiguous bytes; all other tensors keep torch's
// [out][in] layout.  Order = SB3 registration order.
struct Layout {
  int w_pi0, b_pi0, w_pi1, b_pi1, w_vf0, b_vf0, w_vf1, b_vf1, w_act, b_act, w_val, b_val, total;
};

__host__ __device__ inline Layout make_layout(int F, int L) {
  Layout o;
  int p = 0;
  o.w_pi0 = p; p += HID * F;
  o.b_pi0 = p; p += HID;
  o.w_pi1 = p; p += HID * HID;
  o.b_pi1 = p; p += HID;
  o.w_vf0 = p; p += HID * F;
  o.b_vf0 = p; p += HID;
  o.w_vf1 = p; p += HID * HID;
  o.b_vf1 = p; p += HID;
  o.w_act = p; p += L * HID;
  o.b_act = p; p += L;
  o.w_val = p; p += HID;
  o.b_val = p; p += 1;
  o.total = p;
  return o;
}

inline int fill_space(const pth_space* sp, SpaceDev* d) {
  int F = pth_space_feature_dim(sp), L = pth_space_logit_dim(sp);
  if (F < 0 || L < 0 || L > MAXL) return -1;
  memset(d, 0, sizeof(*d));
  d->obs_kind = sp->obs_kind;
  d->obs_len = sp->obs_len;
  d->n_heads = sp->n_heads;
  d->F = F;
  d->L = L;
  for (int h = 0; h < sp->n_heads; ++h) d->head_n[h] = sp->head_n[h];
  if (sp->obs_kind == PTH_OBS_ONEHOT) {
    if (sp->obs_len > PTH_MAX_OBS_SLOTS) return -1;  // obs rows are 32 or 96 bytes
    int off = 0;
    for (int s = 0; s < sp->obs_len; ++s) {
      d->slot_off[s] = (int16_t)off;
      off += sp->obs_nvec[s];
    }
  }
  return 0;
}

// Shared-memory image of everything except the two first-layer matrices.
struct SmemPolicy {
  float w_pi1[HID * LDW];
  float w_vf1[HID * LDW];
  float w_act[MAXL * LDW];
  float w_val[HID];
  float b_pi0[HID], b_pi1[HID], b_vf0[HID], b_vf1[HID];
  float b_act[MAXL];
  float b_val;
  float pad_[3];
};

// COHERENT = true: the parameters are rewritten by other CTAs between grid-wide
// barriers inside one kernel (the update).  One-time loads then go to L2
// (ld.global.cg); the hot first-layer row gathers use ordinary L1-cached loads
// (NOT the non-coherent ld.global.nc path) and rely on the gpu-scope fence the
// update kernel executes after its parameter-update barrier to invalidate L1.
template <bool COHERENT>
__device__ __forceinline__ float ld_param(const float* p) {
  return COHERENT ? __ldcg(p) : __ldg(p);
}
template <bool COHERENT>
__device__ __forceinline__ float4 ld_param4(const float4* p) {
  return COHERENT ? *p : __ldg(p);
}
template <bool COHERENT>
__device__ __forceinline__ float4 ld_param4_l2(const float4* p) {
  return COHERENT ? __ldcg(p) : __ldg(p);
}

template <bool COHERENT = false>
__device__ __forceinline__ void load_policy(SmemPolicy& s, const float* p,
                                            const Layout& lo, int L, int tid, int nthreads) {
  // the three matrices as float4 (rows of 64 floats; every tensor starts at a multiple of 4
  // floats of a 16-byte aligned vector): a handful of independent 16-byte loads per thread
  const float4* pi1 = reinterpret_cast<const float4*>(p + lo.w_pi1);
  const float4* vf1 = reinterpret_cast<const float4*>(p + lo.w_vf1);
  const float4* act = reinterpret_cast<const float4*>(p + lo.w_act);
  for (int i = tid; i < HID * HID / 4; i += nthreads) {
    const int j = i >> 4, k = (i & 15) * 4;
    *reinterpret_cast<float4*>(s.w_pi1 + j * LDW + k) = ld_param4_l2<COHERENT>(pi1 + i);
    *reinterpret_cast<float4*>(s.w_vf1 + j * LDW + k) = ld_param4_l2<COHERENT>(vf1 + i);
  }
  for (int i = tid; i < L * HID / 4; i += nthreads) {
    const int j = i >> 4, k = (i & 15) * 4;
    *reinterpret_cast<float4*>(s.w_act + j * LDW + k) = ld_param4_l2<COHERENT>(act + i);
  }
  for (int i = tid; i < HID; i += nthreads) {
    s.w_val[i] = ld_param<COHERENT>(p + lo.w_val + i);
    s.b_pi0[i] = ld_param<COHERENT>(p + lo.b_pi0 + i);
    s.b_pi1[i] = ld_param<COHERENT>(p + lo.b_pi1 + i);
    s.b_vf0[i] = ld_param<COHERENT>(p + lo.b_vf0 + i);
    s.b_vf1[i] = ld_param<COHERENT>(p + lo.b_vf1 + i);
  }
  for (int i = tid; i < L; i += nthreads) s.b_act[i] = ld_param<COHERENT>(p + lo.b_act + i);
  if (tid == 0) s.b_val = ld_param<COHERENT>(p + lo.b_val);
}

// load_policy for a 512-thread CTA with every load issued before the first store: one L2 round trip
// instead of one per tensor group (the update kernel reloads the weights 320 times per launch).
template <bool COHERENT>
__device__ __forceinline__ void load_policy_512(SmemPolicy& s, const float* p, const Layout& lo, int L, int tid) {
  static_assert(MAXL * (HID / 4) <= 512 && HID * HID / 4 == 2 * 512, "one pass of 512 threads per tensor");
  const float4* pi1 = reinterpret_cast<const float4*>(p + lo.w_pi1);
  const float4* vf1 = reinterpret_cast<const float4*>(p + lo.w_vf1);
  const float4* act = reinterpret_cast<const float4*>(p + lo.w_act);
  float4 m[2][2], a = make_float4(0.f, 0.f, 0.f, 0.f);
  float sc[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, ba = 0.f, bv = 0.f;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    m[0][q] = ld_param4_l2<COHERENT>(pi1 + tid + q * 512);
    m[1][q] = ld_param4_l2<COHERENT>(vf1 + tid + q * 512);
  }
  if (tid < L * (HID / 4)) a = ld_param4_l2<COHERENT>(act + tid);
  if (tid < HID) {
    sc[0] = ld_param<COHERENT>(p + lo.w_val + tid);
    sc[1] = ld_param<COHERENT>(p + lo.b_pi0 + tid);
    sc[2] = ld_param<COHERENT>(p + lo.b_pi1 + tid);
    sc[3] = ld_param<COHERENT>(p + lo.b_vf0 + tid);
    sc[4] = ld_param<COHERENT>(p + lo.b_vf1 + tid);
  }
  if (tid < L) ba = ld_param<COHERENT>(p + lo.b_act + tid);
  if (tid == 0) bv = ld_param<COHERENT>(p + lo.b_val);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = tid + q * 512, j = i >> 4, k = (i & 15) * 4;
    *reinterpret_cast<float4*>(s.w_pi1 + j * LDW + k) = m[0][q];
    *reinterpret_cast<float4*>(s.w_vf1 + j * LDW + k) = m[1][q];
  }
  if (tid < L * (HID / 4)) *reinterpret_cast<float4*>(s.w_act + (tid >> 4) * LDW + (tid & 15) * 4) = a;
  if (tid < HID) {
    s.w_val[tid] = sc[0];
    s.b_pi0[tid] = sc[1];
    s.b_pi1[tid] = sc[2];
    s.b_vf0[tid] = sc[3];
    s.b_vf1[tid] = sc[4];
  }
  if (tid < L) s.b_act[tid] = ba;
  if (tid == 0) s.b_val = bv;
}

// ---------------------------------------------------------------------------
// First layer, one-hot observations: Out[j][b] = tanh(bias[j] + sum_s W[f_s][j]),
// f_s = slot_off[s] + obs[b][s], slots DESCENDING (numeric contract, DESIGN.md 3: the chain then
// begins with the padded trailing slots, which the update kernel evaluates once per tile).  W is input-major [F][64] in
// global memory (L1/L2 resident).  16 threads cover one row with float4 loads,
// so the CTA works on 8 samples at a time; 4 sample groups are interleaved for
// ILP.  obs: [BT][OW] bytes in shared memory (OW = 32, or 96 for spaces with more than 32 slots).
template <bool COHERENT = false, int NTH = NT, int BTS = BT, int SB = 6, int OW = 32>
__device__ __forceinline__ void first_layer_onehot(const SpaceDev& sp, const uint8_t* obs_s,
                                                   const float* W,
                                                   const float* bias_s, float* Out, int tid,
                                                   bool apply_tanh = true) {
  constexpr int LDA = BTS + 4;
  constexpr int GS = NTH / 16;  // samples the CTA covers at a time (16 threads per row)
  constexpr int UI = (BTS / GS) < 4 ? (BTS / GS) : 4;  // sample groups interleaved for ILP
  const int jq = tid & 15;   // outputs jq*4 .. +3
  const int bs = tid >> 4;   // sample within the group of GS
  const float4 bv = *reinterpret_cast<const float4*>(bias_s + jq * 4);
  const float4* W4 = reinterpret_cast<const float4*>(W);
#pragma unroll 1
  for (int g0 = 0; g0 < BTS / GS; g0 += UI) {
    float4 acc[UI];
#pragma unroll
    for (int u = 0; u < UI; ++u) acc[u] = bv;
    // slots in blocks of SB (template: 6 at 255 registers per thread, 3 at 128): all SB*4 row
    // loads are issued before the adds
    // (the adds themselves stay in descending slot order per sample).
    int s0 = sp.obs_len - 1;
    for (; s0 - SB + 1 >= 0; s0 -= SB) {
      float4 w[SB][UI];
#pragma unroll
      for (int i = 0; i < SB; ++i) {
        const int off = sp.slot_off[s0 - i];
#pragma unroll
        for (int u = 0; u < UI; ++u) {
          const int b = (g0 + u) * GS + bs;
          const int f = off + obs_s[b * OW + s0 - i];
          w[i][u] = ld_param4<COHERENT>(W4 + f * (HID / 4) + jq);
        }
      }
#pragma unroll
      for (int i = 0; i < SB; ++i)
#pragma unroll
        for (int u = 0; u < UI; ++u) {
          acc[u].x = acc[u].x + w[i][u].x;
          acc[u].y = acc[u].y + w[i][u].y;
          acc[u].z = acc[u].z + w[i][u].z;
          acc[u].w = acc[u].w + w[i][u].w;
        }
    }
    for (int s = s0; s >= 0; --s) {
      const int off = sp.slot_off[s];
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        const int b = (g0 + u) * GS + bs;
        const int f = off + obs_s[b * OW + s];
        const float4 w = ld_param4<COHERENT>(W4 + f * (HID / 4) + jq);
        acc[u].x = acc[u].x + w.x;
        acc[u].y = acc[u].y + w.y;
        acc[u].z = acc[u].z + w.z;
        acc[u].w = acc[u].w + w.w;
      }
    }
#pragma unroll
    for (int u = 0; u < UI; ++u) {
      const int b = (g0 + u) * GS + bs;
      float* o = Out + (jq * 4) * LDA + b;
      if (apply_tanh) {
        o[0 * LDA] = pth_tanhf(acc[u].x);
        o[1 * LDA] = pth_tanhf(acc[u].y);
        o[2 * LDA] = pth_tanhf(acc[u].z);
        o[3 * LDA] = pth_tanhf(acc[u].w);
      } else {
        o[0 * LDA] = acc[u].x;
        o[1 * LDA] = acc[u].y;
        o[2 * LDA] = acc[u].z;
        o[3 * LDA] = acc[u].w;
      }
    }
  }
}

// tanh applied in place to the NQ float4 a thread has just written itself (no barrier needed),
// as a rolled loop: the polynomial tanh is ~32 instructions, and 32 inlined copies per layer
// made the epilogues instruction-fetch bound (ncu: stall_no_inst on every tanh line).
template <int NQ, typename F>
__device__ __forceinline__ void tanh_own_quads(float* Out, int tid, F index_of) {
  (void)tid;
#pragma unroll 1
  for (int q = 0; q < NQ; ++q) {
    float4* o = reinterpret_cast<float4*>(Out + index_of(q));
    float4 v = *o;
    v.x = pth_tanhf(v.x);
    v.y = pth_tanhf(v.y);
    v.z = pth_tanhf(v.z);
    v.w = pth_tanhf(v.w);
    *o = v;
  }
}

// First layer, Box observations: X[k][b] (k < F) feature-major in shared memory,
// W input-major [F][64] in global memory (L1/L2 resident).  Same thread tile as
// dense64: SPT samples x JT contiguous outputs; Out[j][b] = tanh(bias[j] +
// sum_k fma(X[k][b], W[k][j], .)), k ascending.
template <bool COHERENT = false, int NTH = NT, int BTS = BT, bool TANH = true>
__device__ __forceinline__ void first_layer_box(int F, const float* X, const float* W,
                                                const float* bias_s, float* Out, int tid) {
  constexpr int LDA = BTS + 4;
  constexpr int SPT = BTS >= 128 ? 8 : 4;
  constexpr int TXN = BTS / SPT, NY = NTH / TXN, JT = HID / NY;
  static_assert(NY * JT == HID && TXN * NY == NTH && (JT == 1 || JT == 2 || JT == 4), "tile mapping");
  const int tx = tid % TXN, ty = tid / TXN;
  float acc[JT][SPT];
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) {
    const float bj = bias_s[ty * JT + jj];
#pragma unroll
    for (int ss = 0; ss < SPT; ++ss) acc[jj][ss] = bj;
  }
#pragma unroll 2
  for (int k = 0; k < F; ++k) {
    float a[SPT];
    {
      const float4 a0 = *reinterpret_cast<const float4*>(X + k * LDA + tx * 4);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      if constexpr (SPT == 8) {
        const float4 a1 = *reinterpret_cast<const float4*>(X + k * LDA + 64 + tx * 4);
        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      }
    }
    float w[JT];
    if constexpr (JT == 4) {
      const float4 wq = ld_param4<COHERENT>(reinterpret_cast<const float4*>(W + k * HID + ty * JT));
      w[0] = wq.x; w[1] = wq.y; w[2] = wq.z; w[3] = wq.w;
    } else if constexpr (JT == 2) {
      const float2* q = reinterpret_cast<const float2*>(W + k * HID + ty * JT);
      const float2 wq = COHERENT ? *q : __ldg(q);
      w[0] = wq.x; w[1] = wq.y;
    } else {
      w[0] = ld_param<COHERENT>(W + k * HID + ty);
    }
#pragma unroll
    for (int jj = 0; jj < JT; ++jj)
#pragma unroll
      for (int ss = 0; ss < SPT; ++ss) acc[jj][ss] = fmaf(a[ss], w[jj], acc[jj][ss]);
  }
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) {
    float* o = Out + (ty * JT + jj) * LDA;
    *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[jj][0], acc[jj][1], acc[jj][2], acc[jj][3]);
    if constexpr (SPT == 8)
      *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[jj][4], acc[jj][5], acc[jj][6], acc[jj][7]);
  }
  if constexpr (TANH)
    tanh_own_quads<JT * (SPT / 4)>(Out, tid, [&](int q) { return (ty * JT + q / (SPT / 4)) * LDA + (q % (SPT / 4)) * 64 + tx * 4; });
}

// AdapPolicy (pantheonrl/algos/adap/policies.py:86-106: features = cat(features, context)): the C
// context inputs continue the first layer's chain behind the features, c ascending, then tanh:
// Out[j][b] = tanh(fma(ctx[C-1][b], Wc[C-1][j], ... fma(ctx[0][b], Wc[0][j], Out[j][b]))).
// Out holds the pre-activations (first layer run without tanh); Cx: [C][LDA] context values per
// sample in shared memory; Wc: rows F .. F + C - 1 of the input-major first-layer matrix (global).
// Warp w of the NW warps owns outputs w, w + NW, ...; lane owns samples lane, lane + 32, ...
template <bool COHERENT = false, int BTS = BT>
__device__ __forceinline__ void context_columns_tanh(int C, const float* Cx, const float* Wc, float* Out,
                                                     int warp, int n_warps, int lane) {
  constexpr int LDA = BTS + 4;
  for (int j = warp; j < HID; j += n_warps) {
    float w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) w[c] = c < C ? ld_param<COHERENT>(Wc + c * HID + j) : 0.f;
#pragma unroll 1
    for (int b = lane; b < BTS; b += 32) {
      float z = Out[j * LDA + b];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < C) z = fmaf(Cx[c * LDA + b], w[c], z);
      Out[j * LDA + b] = pth_tanhf(z);
    }
  }
}

// ---------------------------------------------------------------------------
// Hidden layer: Out[j][b] = act(bias[j] + sum_{k<64} A[k][b] * W[j][k]).
// Thread (tx, ty) owns SPT samples (BTS = 128: {tx*4..+3, 64+tx*4..+3}; smaller
// tiles: tx*4..+3) and the JT outputs j = jj*NY + ty.  BTS = 128, NTH = 256: per
// 4 k, 4 LDS.128 of weights + 8 LDS.128 of activations feed 128 FFMA per thread.
// LDAX: row stride of A / Out when the BTS samples are a column block of a wider tile.
template <bool TANH, int NTH = NT, int BTS = BT, int LDAX = BTS + 4>
__device__ __forceinline__ void dense64(const float* A, const float* W, const float* bias,
                                        float* Out, int tid) {
  constexpr int LDA = LDAX;
  constexpr int SPT = BTS >= 128 ? 8 : 4;
  constexpr int TXN = BTS / SPT, NY = NTH / TXN, JT = HID / NY;
  static_assert(NY * JT == HID && TXN * NY == NTH, "tile mapping must cover 64 outputs x BTS samples");
  const int tx = tid % TXN, ty = tid / TXN;
  float acc[JT][SPT];
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) {
    const float bj = bias[jj * NY + ty];
#pragma unroll
    for (int ss = 0; ss < SPT; ++ss) acc[jj][ss] = bj;
  }
#pragma unroll 2
  for (int k0 = 0; k0 < HID; k0 += 4) {
    float4 w[JT];
#pragma unroll
    for (int jj = 0; jj < JT; ++jj)
      w[jj] = *reinterpret_cast<const float4*>(W + (jj * NY + ty) * LDW + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float a[SPT];
      {
        const float4 a0 = *reinterpret_cast<const float4*>(A + (k0 + kk) * LDA + tx * 4);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        if constexpr (SPT == 8) {
          const float4 a1 = *reinterpret_cast<const float4*>(A + (k0 + kk) * LDA + 64 + tx * 4);
          a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        }
      }
#pragma unroll
      for (int jj = 0; jj < JT; ++jj) {
        const float wk = kk == 0 ? w[jj].x : (kk == 1 ? w[jj].y : (kk == 2 ? w[jj].z : w[jj].w));
#pragma unroll
        for (int ss = 0; ss < SPT; ++ss) acc[jj][ss] = fmaf(a[ss], wk, acc[jj][ss]);
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) {
    float* o = Out + (jj * NY + ty) * LDA;
    *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[jj][0], acc[jj][1], acc[jj][2], acc[jj][3]);
    if constexpr (SPT == 8)
      *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[jj][4], acc[jj][5], acc[jj][6], acc[jj][7]);
  }
  if constexpr (TANH)
    tanh_own_quads<JT * (SPT / 4)>(Out, tid, [&](int q) { return ((q / (SPT / 4)) * NY + ty) * LDA + (q % (SPT / 4)) * 64 + tx * 4; });
}

// ---------------------------------------------------------------------------
// Heads, one thread per sample.  The thread pulls its 64 latent activations
// into registers once; weights are broadcast reads.
template <int BTS = BT>
__device__ __forceinline__ void load_column(const float* A, int b, float (&h)[HID]) {
  constexpr int LDA = BTS + 4;
#pragma unroll
  for (int k = 0; k < HID; ++k) h[k] = A[k * LDA + b];
}

__device__ __forceinline__ float dot64(const float (&h)[HID], const float* w, float bias) {
  float acc = bias;
#pragma unroll
  for (int k0 = 0; k0 < HID; k0 += 4) {
    const float4 wv = *reinterpret_cast<const float4*>(w + k0);
    acc = fmaf(h[k0 + 0], wv.x, acc);
    acc = fmaf(h[k0 + 1], wv.y, acc);
    acc = fmaf(h[k0 + 2], wv.z, acc);
    acc = fmaf(h[k0 + 3], wv.w, acc);
  }
  return acc;
}

// logits[l][b] for l < L into Lg (row stride LDA)
template <int BTS = BT>
__device__ __forceinline__ void action_head(const float* A, const SmemPolicy& s, int L, float* Lg,
                                            int b) {
  constexpr int LDA = BTS + 4;
  float h[HID];
  load_column<BTS>(A, b, h);
  for (int l = 0; l < L; ++l) Lg[l * LDA + b] = dot64(h, s.w_act + l * LDW, s.b_act[l]);
}

template <int BTS = BT>
__device__ __forceinline__ float value_head(const float* A, const SmemPolicy& s, int b) {
  float h[HID];
  load_column<BTS>(A, b, h);
  return dot64(h, s.w_val, s.b_val);
}

// One categorical head over logits z[i] = Lg[(off+i)*LDA + b].
// Inverse-CDF sampling on one uniform: p_i = exp(z_i - max), S = sum ascending,
// first i with cumsum_i > u*S (fallback n-1).  log_prob = (z_a - max) - log S,
// entropy = sum fma(-(p_i/S), (z_i - max) - log S, .).
struct HeadOut {
  int action;
  float logp, entropy;
};

template <int BTS = BT>
__device__ __forceinline__ HeadOut head_eval(const float* Lg, int off, int n, int b, bool sample,
                                             float u, int action_in, float* probs_out = nullptr) {
  constexpr int LDA = BTS + 4;
  float m = Lg[off * LDA + b];
  for (int i = 1; i < n; ++i) {
    float z = Lg[(off + i) * LDA + b];
    m = z > m ? z : m;
  }
  float S = 0.f;
  for (int i = 0; i < n; ++i) S = S + pth_expf(Lg[(off + i) * LDA + b] - m);
  const float logS = pth_logf(S);
  HeadOut o;
  o.action = action_in;
  if (sample) {
    const float thr = u * S;
    float cum = 0.f;
    int a = n - 1;
    bool found = false;
    for (int i = 0; i < n; ++i) {
      cum = cum + pth_expf(Lg[(off + i) * LDA + b] - m);
      if (!found && cum > thr) {
        a = i;
        found = true;
      }
    }
    o.action = a;
  }
  o.logp = (Lg[(off + o.action) * LDA + b] - m) - logS;
  float ent = 0.f;
  for (int i = 0; i < n; ++i) {
    const float zi = Lg[(off + i) * LDA + b];
    const float lp = (zi - m) - logS;
    const float pi = pth_expf(zi - m) / S;
    ent = fmaf(-pi, lp, ent);
    if (probs_out) probs_out[(off + i) * LDA + b] = pi;
  }
  o.entropy = ent;
  return o;
}

struct DistOut {
  uint32_t action;  // 4 packed bytes
  float logp, entropy;
};

template <int BTS = BT>
__device__ __forceinline__ DistOut dist_eval(const SpaceDev& sp, const float* Lg, int b,
                                             bool sample, pth_u4 rnd, uint32_t action_in,
                                             float* probs_out = nullptr) {
  DistOut d;
  d.action = 0;
  d.logp = 0.f;
  d.entropy = 0.f;
  int off = 0;
  const uint32_t r[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
  for (int h = 0; h < sp.n_heads; ++h) {
    HeadOut o = head_eval<BTS>(Lg, off, sp.head_n[h], b, sample, pth_u01(r[h]),
                          (int)((action_in >> (8 * h)) & 0xffu), probs_out);
    d.action |= ((uint32_t)o.action & 0xffu) << (8 * h);
    d.logp = d.logp + o.logp;
    d.entropy = d.entropy + o.entropy;
    off += sp.head_n[h];
  }
  return d;
}

}  // namespace pthmlp
