// pth_update.cu — a5: the whole of PPO.train for one learner in ONE persistent
// cooperative kernel: per epoch / minibatch {gather by permutation, forward,
// advantage normalisation, clipped surrogate + value + entropy losses, hand
// written backward, ordered cross-CTA gradient reduction, global-norm clip,
// Adam}, plus the shuffle and buffer-compaction helpers.
//
// Replaces model.train() at pantheonrl/common/agents.py:155 -> SB3 PPO.train
// (restated in-tree at pantheonrl/algos/adap/adap_learn.py:229-347; Adam eps at
// pantheonrl/algos/modular/policies.py:84-88).
//
// Variants of the same kernel (template flags; the reduction, clip and Adam phases are shared):
//   WIDE  one-hot observation rows of up to 96 slots (frame stacks)
//   ADAP  pantheonrl/algos/adap: AdapPolicy context inputs + the context KL loss as extra tiles (adap_learn.py:229-347,
//         adap/util.py:97-131); with MULT the AdapPolicyMult towers (adap/policies.py:134-283)
//   MOD   pantheonrl/algos/modular: ModularPolicy's per-partner modules + the marginal regulariser, one partner phase
//         of ModularAlgorithm.train per launch (modular/learn.py:221-351, modular/policies.py:273-396)
//
// Reduction contract (mirrored by oracle/pth_oracle_update.inc): samples of a
// minibatch are cut into tiles of 128; tile t belongs to CTA t mod G; inside a
// tile every gradient entry is a sequential fma chain over the tile's samples in
// ascending order; a CTA adds its tiles in order; CTAs are added in ascending
// order; cross-lane sums use the fixed 128-lane tree.  Three grid-wide barriers
// per minibatch: partials -> (reduce, norm partials) -> (clip, Adam).
#include <cooperative_groups.h>
#include <math.h>

#include "pth_mlp.cuh"

namespace cg = cooperative_groups;
using namespace pthmlp;

namespace {

// The update CTA runs 512 threads on its 128-sample tile (16 warps = 4 per SM sub-partition,
// <= 128 registers per thread): every phase is latency bound at this occupancy rather than
// throughput bound, so twice the warps of the rollout / forward kernels (UNT = 256) pays even
// though the register tiles get smaller.  Thread tiles only decide WHO computes an output.
constexpr int UNT = 512;
constexpr int LDT = 66;  // stride of the transposed dz1 tile [sample][unit] (even: float2 loads)
constexpr int MAX_SLOTS = PTH_MAX_OBS_SLOTS;  // 96: widest one-hot row
constexpr int MAX_ROWS = 1024;  // first-layer rows (one-hot feature width) supported by the update
constexpr int STAGE_ROWS = 160;  // first-layer rows per tower staged in shared memory per tile

// WIDE: one-hot spaces with more than 32 slots (frame-stacked observations, rows of 96 bytes).  They
// only occur in the host-driven n_envs = 1 flow, so the wide kernel keeps the plain first layer
// (every row gathered per sample) and spends the shared memory on the wider tiles instead of on
// the row stage and the chain beginnings.
// PLAIN: the first layers gather every row per sample from L1 / L2 (no row stage, no chain
// beginnings): the wide spaces, and ADAP (host-driven batches of 64 samples: set-up would not pay).
template <bool WIDE, bool PLAIN = WIDE>
struct alignas(16) UpdSmemT {
  static constexpr int OW = WIDE ? 96 : 32;  // bytes per observation row = slots supported
  SmemPolicy pol;
  float H1[HID * LDA];
  float H2[HID * LDA];  // later: dz1 transposed [sample][LDT]
  float D1[HID * LDA];  // dz2
  float Lg[MAXL * LDA]; // logits -> dlogits -> first-layer gradient chunk
  uint32_t obs[BT * OW / 4];
  // per slot: the tile's samples stably sorted by observed value, and per
  // first-layer row the number of samples selecting it (built once per tile,
  // used by both towers)
  uint8_t order[OW * BT];
  uint8_t rcount[MAX_ROWS];
  float red[8];
  float bc[160];  // broadcast scratch (norm partials)
  // Common beginnings of the first-layer chains (one-hot): the layer adds its rows in descending
  // slot order, and most samples of a tile agree on the trailing slots (padding).  With d_s the
  // MODE of slot s over the tile, pchain[t][j] = bias + row(S-1, d_{S-1}) + ... + row(j, d_j) for
  // tower t; a sample that agrees with the mode on every slot >= jb starts from pchain[t][jb] and
  // adds only its slots jb-1 .. 0 — the same additions in the same order, hence the same bits.
  float pchain[2][PLAIN ? 4 : (32 + 1) * HID];
  uint8_t dmode[OW];    // mode value per slot (0 beyond obs_len)
  uint8_t jb[BT];            // per sample: slots [0, jb) are its own
  // The rows of the two first-layer matrices that this tile's samples select (~92 of Liar's 270)
  // are staged in shared memory once per tile (in H2 | D1 | Lg, free until the hidden layer runs):
  // without it every (sample, slot) pair pulls 256 bytes from L2 and the layer is bound by the
  // L2 -> SM bandwidth (1.9 MB per tile before, 0.5 MB with the chain beginnings, 47 KB staged).
  // rowmap[row] = staged position, 0xFF = not staged (more than STAGE_ROWS rows touched: global load).
  uint8_t rowmap[MAX_ROWS];
  uint16_t staged_rows[STAGE_ROWS];  // inverse of rowmap
  uint8_t rowpos[BT * 32];           // per (sample, slot): staged position of the row it selects (both towers)
  int n_staged;
};
static_assert(STAGE_ROWS * 2 * HID <= (2 * HID + MAXL) * LDA, "stage fits in H2 | D1 | Lg");
static_assert(STAGE_ROWS < 255, "rowmap is one byte per row");

struct UpdParams {
  SpaceDev sp;
  Layout lo;
  float* params;
  float* adam_m;
  float* adam_v;
  const uint8_t* obs;
  const uint8_t* actions;
  const uint8_t* old_logp;
  const uint8_t* adv;
  const uint8_t* ret;
  int64_t obs_stride, act_stride, f_stride;  // bytes
  const int32_t* index;
  const int32_t* perm;
  int64_t M, BS;
  int n_epochs;
  float lr, clip, ent_coef, vf_coef, max_norm, b1, b2, eps;
  int normalize;
  int loss_kind;  // PTH_LOSS_PPO / PTH_LOSS_BC
  float l2;       // BC: weight of sum(theta^2) / 2
  double b1pow0, b2pow0;
  float* part;       // [G][P]
  float* grad;       // [P]
  float* norm_part;  // [G]
  float* stat_part;  // [G][8]
  float* advstat;    // [n_epochs * n_mb][2]
  float* stats;      // [n_epochs * n_mb][8] or NULL
  uint8_t nvec[MAX_SLOTS];
  // multi-GPU sharded update (world > 1)
  int world, rank;
  uint2* xbuf[8];       // rank r's exchange buffer [2 parities][world][XS] of (value bits, epoch tag)
  uint2* advx[8];       // rank r's advantage-statistics exchange [n_epochs * n_mb][2] of (value bits, launch tag)
  uint32_t flag_epoch;
  int XS;               // floats per (parity, rank) slot: P + 8 rounded up to 4
  long long* prof;      // debug: per-phase clock64 sums of CTA 0 (pth_debug_update_profile), or NULL
  // ADAP (pantheonrl/algos/adap): C context inputs behind the features of both first layers, and the
  // context loss of adap/util.py:97-131 as extra tiles of every minibatch
  int C;                     // context_size (0: plain MlpPolicy)
  const float* ctx;          // [rows][C] context stored with every sample
  float ctx_coeff;           // context_loss_coeff; 0: no context tiles
  int K, NS;                 // num_context_samples, num_state_samples
  const int32_t* ctx_sidx;   // [n_epochs * n_mb][NS] positions inside the minibatch
  const float* ctx_draws;    // [n_epochs * n_mb][K][C] sampled contexts
  float* ctx_loss;           // [n_epochs * n_mb] or NULL (MOD: the marginal regulariser of every minibatch)
  // ModularAlgorithm (pantheonrl/algos/modular): Pn partner modules behind the main network
  int P;                     // all parameters (lo.total, + Pn partner blocks when MOD)
  int Pn, p0;                // num_partners, the partner whose buffer this launch trains on
  float marg_coef;           // marginal_reg_coef
  double b1pow0_v, b2pow0_v; // beta^steps of p0's value modules (their own Adam step count)
  float* mod_ws;             // [G] per-CTA scratch of mod_scratch_floats(Pn) (MULT: mult_scratch_floats(C)) floats
  // AdapPolicyMult (adap/policies.py:134-283): offsets of the two scaling layers (64 -> 64 C), tower 0 = policy
  int mult_w[2], mult_b[2];
};
constexpr int MAX_CTX = 8;

// phase timeline of CTA 0 (thread 0), accumulated over every minibatch of the launch
#define PTH_PROF(i)                                   \
  do {                                                \
    if (p.prof != nullptr && c == 0 && tid == 0) {    \
      const long long now__ = clock64();              \
      p.prof[i] += now__ - prof_last;                 \
      prof_last = now__;                              \
    }                                                 \
  } while (0)

// The fixed 128-lane tree of the reduction contract: xor-shuffle tree inside
// each of the first four warps, then the four warp sums left to right.  Threads
// [128, UNT) take part in the barriers only.
__device__ __forceinline__ float block_tree(float x, float* red, int tid) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, d);
  __syncthreads();  // red may still be read by the previous call
  if ((tid & 31) == 0 && tid < BT) red[tid >> 5] = x;
  __syncthreads();
  return ((red[0] + red[1]) + red[2]) + red[3];
}

// Partial-gradient accumulation across a CTA's tiles.  Every address is owned by
// exactly one thread of one CTA, so tile t+1's add follows tile t's in program
// order; a fire-and-forget L2 reduction (RED.ADD.F32, IEEE round-to-nearest)
// gives the same bits as load-add-store without the load round trip.
// Cross-GPU exchange of the per-rank gradient sums, "LL" style: every value travels as ONE
// 8-byte store {value bits, epoch tag} into the peer's exchange buffer (8-byte stores are single
// NVLink transactions), and the reader polls the pair itself until the tag is this minibatch's
// epoch.  No fence, no separate flag, no barrier: a parameter slice only ever meets the same
// slice of the other ranks, and a value is usable the moment it lands.  Buffers are zero
// initialised once and epochs start at 1; two parities keep a fast rank's next-but-one
// minibatch off values a slow rank is still reading (it cannot get that far ahead: it needs the
// slow rank's next values first).
__device__ __forceinline__ void ll_store(uint2* dst, float v, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(epoch)
               : "memory");
}
__device__ __forceinline__ uint2 ll_peek(const uint2* src) {
  uint2 r;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(src) : "memory");
  return r;
}
__device__ __forceinline__ float ll_wait(const uint2* src, uint32_t epoch, uint2 first) {
  long long spins = 0;
  while (first.y != epoch) {
    first = ll_peek(src);
    if (++spins > (1ll << 24)) __trap();  // a peer died: fail loudly instead of hanging the GPU
  }
  return __uint_as_float(first.x);
}

__device__ __forceinline__ void acc_store(float* g, float v, bool first) {
  if (first)
    *g = v;
  else
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g), "f"(v) : "memory");
}

__device__ __forceinline__ int64_t sample_offset(const UpdParams& p, int e, int64_t i) {
  const int64_t j = __ldg(p.perm + (int64_t)e * p.M + i);
  return p.index ? (int64_t)__ldg(p.index + j) : j;
}

// gW[j][k] = sum_b fma(Dz[j][b], Hh[k][b], .)  (64 x 64 outputs, b ascending)
// (ldg: row stride of gout — AdapPolicyMult's scaling sub-layers own every C-th row of their matrix)
template <int NTH>
__device__ __forceinline__ void wgrad64(const float* Dz, const float* Hh, float* gout, bool first,
                                        int tid, int ldg = HID) {
  constexpr int NY = NTH / 16, JT = HID / NY;
  const int kt = tid & 15, jt = tid >> 4;
  float acc[JT][4];
#pragma unroll
  for (int jj = 0; jj < JT; ++jj)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) acc[jj][kk] = 0.f;
#pragma unroll 2
  for (int b0 = 0; b0 < BT; b0 += 4) {
    float4 d[JT], h[4];
#pragma unroll
    for (int jj = 0; jj < JT; ++jj)
      d[jj] = *reinterpret_cast<const float4*>(Dz + (jt + NY * jj) * LDA + b0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      h[kk] = *reinterpret_cast<const float4*>(Hh + (kt + 16 * kk) * LDA + b0);
#pragma unroll
    for (int jj = 0; jj < JT; ++jj)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float a = acc[jj][kk];
        a = fmaf(d[jj].x, h[kk].x, a);
        a = fmaf(d[jj].y, h[kk].y, a);
        a = fmaf(d[jj].z, h[kk].z, a);
        a = fmaf(d[jj].w, h[kk].w, a);
        acc[jj][kk] = a;
      }
  }
#pragma unroll
  for (int jj = 0; jj < JT; ++jj)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      acc_store(gout + (jt + NY * jj) * ldg + (kt + 16 * kk), acc[jj][kk], first);
}

// head weight gradient: gWa[l][k] = sum_b fma(Lg[l][b], H2[k][b], .), l < L
__device__ __forceinline__ void head_wgrad(const float* Lg, const float* Hh, int L, float* gout,
                                           bool first, int tid) {
  const int nlt = (L + 3) >> 2;
  if (tid >= nlt * 16) return;
  const int kt = tid & 15, lt = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int ll = 0; ll < 4; ++ll)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) acc[ll][kk] = 0.f;
  for (int b0 = 0; b0 < BT; b0 += 4) {
    float4 d[4], h[4];
#pragma unroll
    for (int ll = 0; ll < 4; ++ll) {
      const int l = lt * 4 + ll;
      d[ll] = l < L ? *reinterpret_cast<const float4*>(Lg + l * LDA + b0) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      h[kk] = *reinterpret_cast<const float4*>(Hh + (kt + 16 * kk) * LDA + b0);
#pragma unroll
    for (int ll = 0; ll < 4; ++ll)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float a = acc[ll][kk];
        a = fmaf(d[ll].x, h[kk].x, a);
        a = fmaf(d[ll].y, h[kk].y, a);
        a = fmaf(d[ll].z, h[kk].z, a);
        a = fmaf(d[ll].w, h[kk].w, a);
        acc[ll][kk] = a;
      }
  }
#pragma unroll
  for (int ll = 0; ll < 4; ++ll) {
    const int l = lt * 4 + ll;
    if (l < L)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) acc_store(gout + l * HID + (kt + 16 * kk), acc[ll][kk], first);
  }
}

// row sums over the tile: gout[r] = sum_b X[r][b], threads [t0, t0 + rows)
__device__ __forceinline__ void row_sums(const float* X, int rows, float* gout, bool first, int tid,
                                         int t0, int gstride = 1) {
  const int r = tid - t0;
  if (r < 0 || r >= rows) return;
  float s = 0.f;
  for (int b0 = 0; b0 < BT; b0 += 4) {
    const float4 v = *reinterpret_cast<const float4*>(X + r * LDA + b0);
    s = s + v.x;
    s = s + v.y;
    s = s + v.z;
    s = s + v.w;
  }
  acc_store(gout + r * gstride, s, first);
}

// dz1[k][b] = (sum_j fma(W[j][k], Dz[j][b], .)) * (1 - Hact[k][b]^2), stored transposed
// ([b][LDT], one-hot path: the segmented sums walk samples) or feature-major ([k][LDA], Box path)
template <bool TRANSPOSED, int NTH>
__device__ __forceinline__ void backprop64(const float* Dz, const float* W, const float* Hact,
                                           float* outT, int tid) {
  constexpr int KT = HID / (NTH / 16);  // outputs per thread, contiguous
  const int tx = tid & 15, ty = tid >> 4;
  float acc[KT][8];
#pragma unroll
  for (int kk = 0; kk < KT; ++kk)
#pragma unroll
    for (int ss = 0; ss < 8; ++ss) acc[kk][ss] = 0.f;
#pragma unroll 4
  for (int j = 0; j < HID; ++j) {
    const float4 a0 = *reinterpret_cast<const float4*>(Dz + j * LDA + tx * 4);
    const float4 a1 = *reinterpret_cast<const float4*>(Dz + j * LDA + 64 + tx * 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float w[KT];
    if constexpr (KT == 2) {
      const float2 wq = *reinterpret_cast<const float2*>(W + j * LDW + ty * KT);
      w[0] = wq.x; w[1] = wq.y;
    } else {
#pragma unroll
      for (int q = 0; q < KT / 4; ++q) {
        const float4 wq = *reinterpret_cast<const float4*>(W + j * LDW + ty * KT + 4 * q);
        w[4 * q + 0] = wq.x; w[4 * q + 1] = wq.y; w[4 * q + 2] = wq.z; w[4 * q + 3] = wq.w;
      }
    }
#pragma unroll
    for (int kk = 0; kk < KT; ++kk)
#pragma unroll
      for (int ss = 0; ss < 8; ++ss) acc[kk][ss] = fmaf(w[kk], a[ss], acc[kk][ss]);
  }
#pragma unroll
  for (int ss = 0; ss < 8; ++ss) {
    const int b = ss < 4 ? tx * 4 + ss : 64 + tx * 4 + (ss - 4);
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const int k = ty * KT + kk;
      const float h = Hact[k * LDA + b];
      outT[TRANSPOSED ? (b * LDT + k) : (k * LDA + b)] = acc[kk][ss] * (1.0f - h * h);
    }
  }
}

// Box first layer: gW0[k][j] = sum_b fma(Dz1[j][b], X[k][b], .) for k < F, b ascending; the
// tile of wgrad64 with the result stored input-major (the layout of the first-layer matrices).
__device__ __forceinline__ void wgrad_first_box(const float* Dz, const float* X, int F, float* gout,
                                                bool first, int tid) {
  constexpr int NY = UNT / 16, JT = HID / NY;
  const int kt = tid & 15, jt = tid >> 4;
  float acc[JT][4];
#pragma unroll
  for (int jj = 0; jj < JT; ++jj)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) acc[jj][kk] = 0.f;
#pragma unroll 2
  for (int b0 = 0; b0 < BT; b0 += 4) {
    float4 d[JT], h[4];
#pragma unroll
    for (int jj = 0; jj < JT; ++jj)
      d[jj] = *reinterpret_cast<const float4*>(Dz + (jt + NY * jj) * LDA + b0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      h[kk] = *reinterpret_cast<const float4*>(X + (kt + 16 * kk) * LDA + b0);
#pragma unroll
    for (int jj = 0; jj < JT; ++jj)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float a = acc[jj][kk];
        a = fmaf(d[jj].x, h[kk].x, a);
        a = fmaf(d[jj].y, h[kk].y, a);
        a = fmaf(d[jj].z, h[kk].z, a);
        a = fmaf(d[jj].w, h[kk].w, a);
        acc[jj][kk] = a;
      }
  }
#pragma unroll
  for (int jj = 0; jj < JT; ++jj)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      if (kt + 16 * kk < F) acc_store(gout + (kt + 16 * kk) * HID + (jt + NY * jj), acc[jj][kk], first);
}

// action head for the whole tile with all UNT threads: Lg[l][b] = b_act[l] + sum_k fma(H2[k][b],
// w_act[l][k], .), k ascending (the chain of dot64).  Warp w owns logits w, w + NW, ... (uniform
// per warp: weight reads are broadcasts); lane owns 4 consecutive samples.
__device__ __forceinline__ void logits_tile_w(const float* Hh, const float* w_act, const float* b_act, int L,
                                              float* Lg, int tid) {
  constexpr int NW = UNT / 32, NL = MAXL / NW;  // warps, logits per warp
  const int tx = tid & 31, ty = tid >> 5;
  float acc[NL][4];
#pragma unroll
  for (int ll = 0; ll < NL; ++ll) {
    const int l = ty + NW * ll;
    const float bl = l < L ? b_act[l] : 0.f;
#pragma unroll
    for (int ss = 0; ss < 4; ++ss) acc[ll][ss] = bl;
  }
#pragma unroll 4
  for (int k0 = 0; k0 < HID; k0 += 4) {
    float4 w[NL];
#pragma unroll
    for (int ll = 0; ll < NL; ++ll) {
      const int l = ty + NW * ll;
      w[ll] = l < L ? *reinterpret_cast<const float4*>(w_act + l * LDW + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(Hh + (k0 + kk) * LDA + tx * 4);
#pragma unroll
      for (int ll = 0; ll < NL; ++ll) {
        const float wk = kk == 0 ? w[ll].x : (kk == 1 ? w[ll].y : (kk == 2 ? w[ll].z : w[ll].w));
        acc[ll][0] = fmaf(a.x, wk, acc[ll][0]);
        acc[ll][1] = fmaf(a.y, wk, acc[ll][1]);
        acc[ll][2] = fmaf(a.z, wk, acc[ll][2]);
        acc[ll][3] = fmaf(a.w, wk, acc[ll][3]);
      }
    }
  }
#pragma unroll
  for (int ll = 0; ll < NL; ++ll) {
    const int l = ty + NW * ll;
    if (l < L)
      *reinterpret_cast<float4*>(Lg + l * LDA + tx * 4) = make_float4(acc[ll][0], acc[ll][1], acc[ll][2], acc[ll][3]);
  }
}

__device__ __forceinline__ void logits_tile(const float* Hh, const SmemPolicy& pol, int L, float* Lg, int tid) {
  logits_tile_w(Hh, pol.w_act, pol.b_act, L, Lg, tid);
}

// ---- ModularPolicy building blocks (pantheonrl/algos/modular/policies.py) -------------------------
// out[k][b] = (sum_l fma(w_act[l][k], Dl[l][b], .)) * (ACT ? 1 - Hact[k][b]^2 : 1), l ascending: a head's
// share of d loss / d latent.  Four threads per sample, 16 hidden units each; w_act in shared memory.
template <bool ACT>
__device__ __forceinline__ void head_backprop(const float* Dl, const float* w_act, const float* Hact, int L,
                                              float* Out, int tid) {
  static_assert(UNT / BT == 4, "four threads per sample");
  const int sb = tid & (BT - 1), kb = (tid >> 7) * (HID / 4);
  float acc[HID / 4];
#pragma unroll
  for (int k = 0; k < HID / 4; ++k) acc[k] = 0.f;
  for (int l = 0; l < L; ++l) {
    const float d = Dl[l * LDA + sb];
#pragma unroll
    for (int k0 = 0; k0 < HID / 4; k0 += 4) {
      const float4 w = *reinterpret_cast<const float4*>(w_act + l * LDW + kb + k0);
      acc[k0 + 0] = fmaf(w.x, d, acc[k0 + 0]);
      acc[k0 + 1] = fmaf(w.y, d, acc[k0 + 1]);
      acc[k0 + 2] = fmaf(w.z, d, acc[k0 + 2]);
      acc[k0 + 3] = fmaf(w.w, d, acc[k0 + 3]);
    }
  }
#pragma unroll
  for (int k = 0; k < HID / 4; ++k) {
    float v = acc[k];
    if constexpr (ACT) {
      const float h = Hact[(kb + k) * LDA + sb];
      v = v * (1.0f - h * h);
    }
    Out[(kb + k) * LDA + sb] = v;
  }
}
// a [rows][64] matrix from global memory into a shared-memory slot with row stride LDW
// (the partner blocks follow the main network without padding: their matrices need not be 16-byte aligned)
__device__ __forceinline__ void stage_rows64(float* dst, const float* src, int rows, int tid) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = tid; i < rows * (HID / 4); i += UNT) {
      const int j = i >> 4, k = (i & 15) * 4;
      *reinterpret_cast<float4*>(dst + j * LDW + k) = __ldcg(s4 + i);
    }
  } else {
    for (int i = tid; i < rows * HID; i += UNT) dst[(i >> 6) * LDW + (i & 63)] = __ldcg(src + i);
  }
}
__device__ __forceinline__ void stage_vec(float* dst, const float* src, int n, int tid) {
  for (int i = tid; i < n; i += UNT) dst[i] = __ldcg(src + i);
}
// every `rstride`-th row of a [.][64] matrix (AdapPolicyMult: sub-layer c of a scaling layer owns rows j C + c)
__device__ __forceinline__ void stage_rows64_strided(float* dst, const float* src, int rstride, int tid) {
  for (int i = tid; i < HID * HID; i += UNT) dst[(i >> 6) * LDW + (i & 63)] = __ldcg(src + (size_t)(i >> 6) * rstride * HID + (i & 63));
}
__device__ __forceinline__ void stage_vec_strided(float* dst, const float* src, int n, int stride, int tid) {
  for (int i = tid; i < n; i += UNT) dst[i] = __ldcg(src + (size_t)i * stride);
}
// per-CTA global scratch of the AdapPolicyMult tile: activations of both towers kept for the backward pass
struct MultScratch {
  float *x[2], *y[2], *h2[2], *s[2];  // s[t]: [C] tiles
  float *dy, *dx, *tmp, *zero;
};
__host__ __device__ inline size_t mult_scratch_floats(int C) { return (size_t)(10 + 2 * C) * (HID * LDA); }
// per-partner parameter block behind the MlpPolicy layout (oracle/pth_oracle_modular.inc: mod_block)
struct ModBlock {
  int w_pi0, b_pi0, w_pi1, b_pi1, w_vf0, b_vf0, w_vf1, b_vf1, w_act, b_act, w_val, b_val, total;
};
__host__ __device__ inline ModBlock mod_block(int L) {
  ModBlock o;
  int p = 0;
  o.w_pi0 = p; p += HID * HID;
  o.b_pi0 = p; p += HID;
  o.w_pi1 = p; p += HID * HID;
  o.b_pi1 = p; p += HID;
  o.w_vf0 = p; p += HID * HID;
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
constexpr int MOD_MAX_PARTNERS = 8;
constexpr int TILE_F = HID * LDA;    // floats of a [64][LDA] scratch tile
__device__ __forceinline__ MultScratch mult_scratch(float* base, int C) {
  MultScratch g;
  float* q = base;
  for (int t = 0; t < 2; ++t) {
    g.x[t] = q; q += TILE_F;
    g.y[t] = q; q += TILE_F;
    g.h2[t] = q; q += TILE_F;
    g.s[t] = q; q += (size_t)C * TILE_F;
  }
  g.dy = q; q += TILE_F;
  g.dx = q; q += TILE_F;
  g.tmp = q; q += TILE_F;
  g.zero = q;
  return g;
}
constexpr int LTILE_F = MAXL * LDA;  // floats of a [32][LDA] logits-sized scratch tile
// per-CTA global scratch of the modular tile (activations of every module are kept for the backward pass)
struct ModScratch {
  float *a1p, *a1v, *h2, *v2, *r1, *r2, *tmp, *zero, *q1, *q2;  // q1 / q2: [Pn] tiles
  float *lgm, *pm, *dlm, *lgp, *pc, *dlp;                        // logits-sized; lgp / pc / dlp: [Pn]
};
__host__ __device__ inline size_t mod_scratch_floats(int Pn) {
  return (size_t)(8 + 2 * Pn) * TILE_F + (size_t)(3 + 3 * Pn) * LTILE_F;
}
__device__ __forceinline__ ModScratch mod_scratch(float* base, int Pn) {
  ModScratch g;
  float* q = base;
  g.a1p = q; q += TILE_F;
  g.a1v = q; q += TILE_F;
  g.h2 = q; q += TILE_F;
  g.v2 = q; q += TILE_F;
  g.r1 = q; q += TILE_F;
  g.r2 = q; q += TILE_F;
  g.tmp = q; q += TILE_F;
  g.zero = q; q += TILE_F;
  g.q1 = q; q += (size_t)Pn * TILE_F;
  g.q2 = q; q += (size_t)Pn * TILE_F;
  g.lgm = q; q += LTILE_F;
  g.pm = q; q += LTILE_F;
  g.dlm = q; q += LTILE_F;
  g.lgp = q; q += (size_t)Pn * LTILE_F;
  g.pc = q; q += (size_t)Pn * LTILE_F;
  g.dlp = q;
  return g;
}

// Adam group of a parameter under ModularAlgorithm.train: 0 trained in every phase, 1 the value modules of
// the partner being trained (their own step count), 2 the value modules of the other partners (no gradient:
// Adam skips them, torch >= 2 zero_grad semantics)
__device__ __forceinline__ int mod_group(int pi, int main_total, int blk_total, int w_vf0, int w_act, int w_val,
                                         int p0) {
  if (pi < main_total) return 0;
  const int q = (pi - main_total) / blk_total, r = (pi - main_total) - q * blk_total;
  if ((r >= w_vf0 && r < w_act) || r >= w_val) return q == p0 ? 1 : 2;
  return 0;
}

// Stable counting sort of the tile's nb samples by observed value, one slot per warp at a time:
// order[s][pos] = sample id, rcount[row] = samples selecting the first-layer row (row = slot_off[s]
// + value), dmode[s] = the slot's mode (ties: the smallest value), and every selected row gets a
// position in the shared-memory stage (rowmap / staged_rows; positions are handed out with one
// shared-memory atomic per slot — WHICH position a row gets changes no result).
// nvec <= 32 (every space of the built-in games): one MATCH.ANY per 32 samples groups equal values,
// lane v owns value v's count, a warp scan turns counts into start positions.  Larger nvec: one
// ballot per value.  Per-warp scratch (64 ints) lives in sm.rowpos, which is written after the sort.
template <class SM>
__device__ __forceinline__ void sort_slots(const UpdParams& p, SM& sm, const uint8_t* obs_s, int nb,
                                           int lane, int wid, int n_warps) {
  constexpr int OW = SM::OW;
  int* hist = reinterpret_cast<int*>(sm.rowpos) + wid * 64;
  int* cur = hist + 32;
  const unsigned lt = (1u << lane) - 1u;
  uint8_t* order = sm.order;
  for (int s = wid; s < p.sp.obs_len; s += n_warps) {
    int val[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int b = lane + 32 * r;
      val[r] = b < nb ? (int)obs_s[b * OW + s] : -1;
    }
    const int nv = p.nvec[s];
    const int row0 = p.sp.slot_off[s];
    int mode = 0;
    if (nv <= 32) {
      hist[lane] = 0;
      __syncwarp();
      unsigned m[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) m[r] = __match_any_sync(0xffffffffu, val[r]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (val[r] >= 0 && (m[r] & lt) == 0u) hist[val[r]] += __popc(m[r]);  // the group's lowest lane
        __syncwarp();
      }
      const int cnt = hist[lane];
      int x = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
      }
      cur[lane] = x - cnt;  // first position of value `lane`
      if (lane < nv) sm.rcount[row0 + lane] = (uint8_t)cnt;
      mode = 63 - (__reduce_max_sync(0xffffffffu, lane < nv ? cnt * 64 + (63 - lane) : -1) & 63);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (val[r] >= 0) order[s * BT + cur[val[r]] + __popc(m[r] & lt)] = (uint8_t)(lane + 32 * r);
        __syncwarp();
        if (val[r] >= 0 && (m[r] & lt) == 0u) cur[val[r]] += __popc(m[r]);
        __syncwarp();
      }
    } else {
      int base = 0, best = -1;
      for (int v = 0; v < nv; ++v) {
        const int before = base;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const unsigned mm = __ballot_sync(0xffffffffu, val[r] == v);
          if (val[r] == v) order[s * BT + base + __popc(mm & lt)] = (uint8_t)(lane + 32 * r);
          base += __popc(mm);
        }
        if (base - before > best) {  // ties: the smallest value
          best = base - before;
          mode = v;
        }
        if (lane == 0) sm.rcount[row0 + v] = (uint8_t)(base - before);
      }
      __syncwarp();
    }
    if (lane == 0) sm.dmode[s] = (uint8_t)mode;
    // stage positions of the rows this slot's samples select
    for (int v0 = 0; v0 < nv; v0 += 32) {
      const int v = v0 + lane;
      const bool hit = v < nv && sm.rcount[row0 + v] > 0;
      const unsigned t = __ballot_sync(0xffffffffu, hit);
      int base = 0;
      if (lane == 0) base = atomicAdd(&sm.n_staged, __popc(t));
      base = __shfl_sync(0xffffffffu, base, 0);
      const int pos = base + __popc(t & lt);
      const bool staged = hit && pos < STAGE_ROWS;
      if (v < nv) sm.rowmap[row0 + v] = staged ? (uint8_t)pos : (uint8_t)0xFF;
      if (staged) sm.staged_rows[pos] = (uint16_t)(row0 + v);
    }
    __syncwarp();
  }
}

// Per-tile set-up of the one-hot first layers, after sort_slots and a barrier, in ONE step:
//  * the last warp builds both towers' common chain beginnings pchain[t][j], j = S .. 0 (lane = tower
//    x 16 float4 column groups; the mode rows come straight from L2, ten in flight);
//  * the other warps find every sample's jb (its slots [jb, S) agree with the mode), copy the
//    numbered rows of both matrices into the stage (all loads of a thread in flight together) and
//    translate every (sample, slot) into its row's staged position, once for both towers (rowpos).
constexpr int NTC = UNT - 32;  // threads of the copy part
template <class SM>
__device__ __forceinline__ void chain_setup(const UpdParams& p, SM& sm, float* stage, int nb, int tid) {
  const int S = p.sp.obs_len;
  if (tid >= NTC) {
    const int lane = tid & 31;
    const int t = lane >> 4, jq = lane & 15;
    const float4* W4 = reinterpret_cast<const float4*>(p.params + (t ? p.lo.w_vf0 : p.lo.w_pi0));
    float4 acc = *reinterpret_cast<const float4*>((t ? sm.pol.b_vf0 : sm.pol.b_pi0) + jq * 4);
    float* P = sm.pchain[t];
    *reinterpret_cast<float4*>(P + S * HID + jq * 4) = acc;
    constexpr int PB = 10;  // row loads issued together
    for (int s0 = S - 1; s0 >= 0; s0 -= PB) {
      float4 w[PB];
#pragma unroll
      for (int i = 0; i < PB; ++i) {
        const int sl = max(s0 - i, 0);  // clamped: loaded anyway, added only if s0 - i >= 0
        w[i] = ld_param4_l2<true>(W4 + (p.sp.slot_off[sl] + sm.dmode[sl]) * (HID / 4) + jq);
      }
#pragma unroll
      for (int i = 0; i < PB; ++i) {
        const int sl = s0 - i;
        if (sl >= 0) {
          acc.x = acc.x + w[i].x;
          acc.y = acc.y + w[i].y;
          acc.z = acc.z + w[i].z;
          acc.w = acc.w + w[i].w;
          *reinterpret_cast<float4*>(P + sl * HID + jq * 4) = acc;
        }
      }
    }
    return;
  }
  if (tid < BT) {
    const uint32_t* dm = reinterpret_cast<const uint32_t*>(sm.dmode);
    int j = 0;
    if (tid < nb) {
#pragma unroll
      for (int w = 7; w >= 0; --w) {
        const uint32_t x = sm.obs[tid * 8 + w] ^ dm[w];
        if (j == 0 && x != 0u) j = 4 * w + 4 - (__clz((int)x) >> 3);
      }
    }
    sm.jb[tid] = (uint8_t)j;
  }
  const float4* Wp = reinterpret_cast<const float4*>(p.params + p.lo.w_pi0);
  const float4* Wv = reinterpret_cast<const float4*>(p.params + p.lo.w_vf0);
  const int n = (sm.n_staged < STAGE_ROWS ? sm.n_staged : STAGE_ROWS) * 32;  // (staged row, tower, float4 group)
  constexpr int CB = 4;  // loads in flight per thread
  for (int i0 = tid; i0 < n; i0 += CB * NTC) {
    float4 v[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const int i = min(i0 + c * NTC, n - 1);  // clamped: the load is unconditional, the store is not
      const int m = i >> 5, t = (i >> 4) & 1, jq = i & 15;
      v[c] = ld_param4_l2<true>((t ? Wv : Wp) + (int)sm.staged_rows[m] * (HID / 4) + jq);
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const int i = i0 + c * NTC;
      if (i < n) {
        const int m = i >> 5, t = (i >> 4) & 1, jq = i & 15;
        *reinterpret_cast<float4*>(stage + ((size_t)(t * STAGE_ROWS + m)) * HID + jq * 4) = v[c];
      }
    }
  }
  const uint8_t* obs_s = reinterpret_cast<const uint8_t*>(sm.obs);
  for (int i = tid; i < BT * 32; i += NTC) {
    const int sl = i & 31;
    if (sl < S) sm.rowpos[i] = sm.rowmap[p.sp.slot_off[sl] + obs_s[i]];
  }
}
// row of tower t at staged position m (m == 0xFF: not staged, read it from global memory)
__device__ __forceinline__ float4 row_load(const float* stage, const float4* W4, int t, int m, int row, int jq) {
  if (m != 0xFF) return *(reinterpret_cast<const float4*>(stage + ((size_t)(t * STAGE_ROWS + m)) * HID) + jq);
  return ld_param4<true>(W4 + row * (HID / 4) + jq);
}
// First layer of one tower for the whole tile on NTH threads: Out[j][b] = tanh(bias[j] + rows in
// descending slot order).  16 threads cover a row with float4 loads (from the stage); a thread
// group works on 4 samples at a time.  All four start from the chain beginning of the quartet's
// LARGEST jb: a sample with a smaller jb agrees with the mode on the slots in between, so its
// own rows there ARE the mode's rows — no predication, same additions.
template <int NTH, class SM>
__device__ __forceinline__ void first_layer_chain(const UpdParams& p, const SM& sm, const uint8_t* obs_s,
                                                  const float* W, const float* stage, int t, const float* P,
                                                  float* Out, int tid) {
  static_assert(NTH / 16 * 8 == BT, "16 thread groups x 2 quartets cover the tile");
  const int jq = tid & 15, bs = tid >> 4;
  const float4* W4 = reinterpret_cast<const float4*>(W);
  auto row = [&](int b, int sl) {
    const int m = sm.rowpos[b * 32 + sl];
    return row_load(stage, W4, t, m, m != 0xFF ? 0 : p.sp.slot_off[sl] + obs_s[b * 32 + sl], jq);
  };
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int b0 = half * (BT / 2) + bs * 4;  // samples b0 .. b0 + 3
    const uchar4 jq4 = *reinterpret_cast<const uchar4*>(sm.jb + b0);
    int jmax = jq4.x > jq4.y ? jq4.x : jq4.y;
    jmax = jq4.z > jmax ? jq4.z : jmax;
    jmax = jq4.w > jmax ? jq4.w : jmax;
    float4 acc[4];
    acc[0] = *reinterpret_cast<const float4*>(P + jmax * HID + jq * 4);
#pragma unroll
    for (int u = 1; u < 4; ++u) acc[u] = acc[0];
    constexpr int SB = 3;  // slots per block: all SB * 4 row loads are issued before the adds
    int s0 = jmax - 1;
    for (; s0 - SB + 1 >= 0; s0 -= SB) {
      float4 w[SB][4];
#pragma unroll
      for (int i = 0; i < SB; ++i)
#pragma unroll
        for (int u = 0; u < 4; ++u) w[i][u] = row(b0 + u, s0 - i);
#pragma unroll
      for (int i = 0; i < SB; ++i)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[u].x = acc[u].x + w[i][u].x;
          acc[u].y = acc[u].y + w[i][u].y;
          acc[u].z = acc[u].z + w[i][u].z;
          acc[u].w = acc[u].w + w[i][u].w;
        }
    }
    for (; s0 >= 0; --s0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 w = row(b0 + u, s0);
        acc[u].x = acc[u].x + w.x;
        acc[u].y = acc[u].y + w.y;
        acc[u].z = acc[u].z + w.z;
        acc[u].w = acc[u].w + w.w;
      }
    }
    // the thread owns 4 consecutive samples x 4 outputs: one float4 per output row
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float v[4] = {c == 0 ? acc[0].x : c == 1 ? acc[0].y : c == 2 ? acc[0].z : acc[0].w,
                          c == 0 ? acc[1].x : c == 1 ? acc[1].y : c == 2 ? acc[1].z : acc[1].w,
                          c == 0 ? acc[2].x : c == 1 ? acc[2].y : c == 2 ? acc[2].z : acc[2].w,
                          c == 0 ? acc[3].x : c == 1 ? acc[3].y : c == 2 ? acc[3].z : acc[3].w};
      *reinterpret_cast<float4*>(Out + (jq * 4 + c) * LDA + b0) = make_float4(v[0], v[1], v[2], v[3]);
    }
    // tanh in place on what this thread has just written, as a rolled loop (32 inlined copies of
    // the polynomial made the layer instruction-fetch bound, profiles/update_r01b.md)
    tanh_own_quads<4>(Out, tid, [&](int q) { return (jq * 4 + q) * LDA + b0; });
  }
}

// first-layer (one-hot) weight gradient, one slot per warp at a time, lane = column pair.  Per slot
// the row of every value EXCEPT the mode is the ascending register chain over the samples selecting
// it (stable sort: no read-modify-write through memory); the mode's row is the tile's first-layer
// bias gradient (`bsum`, the chain over all samples) minus the sum of the slot's other rows in
// ascending value order (the reduction contract, oracle/pth_oracle_update.inc).  Only the samples
// that differ from the mode are walked: in these games ~17 % of the (slot, sample) pairs.
// Slots are shared out over warps [0, NWS); the caller's barrier separates the walk (phase 0:
// rows of the other values, their sum kept in registers) from the mode rows (phase 1: bsum ready).
constexpr int NWS = UNT / 32 - 2;  // the last two warps compute the bias chains meanwhile
// Two values of the slot at a time, one per half warp (16 lanes x 4 columns: column pairs hl * 2 and
// 32 + hl * 2): the chains of a slot's values are independent, so the slot's critical path halves.
template <class SM, int SLOTS_PER_WARP>
__device__ __forceinline__ void segsum_w1_walk(const UpdParams& p, const SM& sm, const float* dzT,
                                               float* gW0, bool first, int tid, float4 (&csum)[SLOTS_PER_WARP]) {
  const int lane = tid & 31, wid = tid >> 5;
  const int half = lane >> 4, jp0 = (lane & 15) * 2, jp1 = 32 + jp0;
#pragma unroll
  for (int q = 0; q < SLOTS_PER_WARP; ++q) {
    csum[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int s = wid + q * NWS;
    if (wid >= NWS || s >= p.sp.obs_len) continue;
    const int row0 = p.sp.slot_off[s];
    const int nv = p.nvec[s];
    const int mode = sm.dmode[s];
    const uint8_t* ord = sm.order + s * BT;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    int v = 0, pos = 0;
    while (v < nv) {
      if (v == mode) {
        pos += sm.rcount[row0 + v];
        ++v;
        continue;
      }
      // the next two values that are not the mode: A for lanes [0, 16), B (if any) for lanes [16, 32)
      const int vA = v, cntA = sm.rcount[row0 + vA], posA = pos;
      pos += cntA;
      ++v;
      if (v < nv && v == mode) {
        pos += sm.rcount[row0 + v];
        ++v;
      }
      const bool hasB = v < nv;
      const int vB = hasB ? v : vA, cntB = hasB ? sm.rcount[row0 + vB] : 0, posB = pos;
      if (hasB) {
        pos += cntB;
        ++v;
      }
      const int myv = half ? vB : vA, cnt = half ? cntB : cntA, p0 = half ? posB : posA;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int i = 0;
      for (; i + 4 <= cnt; i += 4) {
        float2 d[4], e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* row = dzT + (int)ord[p0 + i + u] * LDT;
          d[u] = *reinterpret_cast<const float2*>(row + jp0);
          e[u] = *reinterpret_cast<const float2*>(row + jp1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a0 = a0 + d[u].x;
          a1 = a1 + d[u].y;
          a2 = a2 + e[u].x;
          a3 = a3 + e[u].y;
        }
      }
      for (; i < cnt; ++i) {
        const float* row = dzT + (int)ord[p0 + i] * LDT;
        const float2 d = *reinterpret_cast<const float2*>(row + jp0);
        const float2 e = *reinterpret_cast<const float2*>(row + jp1);
        a0 = a0 + d.x;
        a1 = a1 + d.y;
        a2 = a2 + e.x;
        a3 = a3 + e.y;
      }
      if (half == 0 || hasB) {
        float* g = gW0 + (row0 + myv) * HID;
        if (first) {
          *reinterpret_cast<float2*>(g + jp0) = make_float2(a0, a1);
          *reinterpret_cast<float2*>(g + jp1) = make_float2(a2, a3);
        } else if (cnt > 0) {
          acc_store(g + jp0, a0, false);
          acc_store(g + jp0 + 1, a1, false);
          acc_store(g + jp1, a2, false);
          acc_store(g + jp1 + 1, a3, false);
        }
      }
      // the slot's running sum of rows, in ascending value order: + A, then + B (an absent B is +0)
      const float o0 = __shfl_xor_sync(0xffffffffu, a0, 16), o1 = __shfl_xor_sync(0xffffffffu, a1, 16);
      const float o2 = __shfl_xor_sync(0xffffffffu, a2, 16), o3 = __shfl_xor_sync(0xffffffffu, a3, 16);
      c.x = (c.x + (half ? o0 : a0)) + (half ? a0 : o0);
      c.y = (c.y + (half ? o1 : a1)) + (half ? a1 : o1);
      c.z = (c.z + (half ? o2 : a2)) + (half ? a2 : o2);
      c.w = (c.w + (half ? o3 : a3)) + (half ? a3 : o3);
    }
    csum[q] = c;
  }
}
template <class SM, int SLOTS_PER_WARP>
__device__ __forceinline__ void segsum_w1_mode(const UpdParams& p, const SM& sm, const float* bsum,
                                               float* gW0, bool first, int tid,
                                               const float4 (&csum)[SLOTS_PER_WARP]) {
  const int lane = tid & 31, wid = tid >> 5;
  const int half = lane >> 4, jp = (lane & 15) * 2 + 32 * half;  // half 0: columns < 32, half 1: the rest
  const float2 b2 = *reinterpret_cast<const float2*>(bsum + jp);
#pragma unroll
  for (int q = 0; q < SLOTS_PER_WARP; ++q) {
    const int s = wid + q * NWS;
    if (wid >= NWS || s >= p.sp.obs_len) continue;
    float* g = gW0 + (p.sp.slot_off[s] + sm.dmode[s]) * HID + jp;
    const float g0 = b2.x - (half ? csum[q].z : csum[q].x), g1 = b2.y - (half ? csum[q].w : csum[q].y);
    if (first) {
      *reinterpret_cast<float2*>(g) = make_float2(g0, g1);
    } else {
      acc_store(g, g0, false);
      acc_store(g + 1, g1, false);
    }
  }
}

// body of one tower after dz2 (in sm.D1) is known; Ha = the tower's first-layer activations.
// The hidden-layer weight gradient (D1 x Ha^T) and the back-propagation through the hidden layer
// (W^T x D1) only READ D1 / Ha, so they run side by side: each on one half of the CTA with the
// large register tiles of a 256-thread group (the 512-thread tiles are shared-memory bound).
// first-layer gradients of a tower from dz1 in sm.H2 ([sample][LDT] for one-hot spaces, [unit][LDA] for Box)
template <bool BOX, bool ADAP = false, class SM>
__device__ __forceinline__ void first_layer_backward(const UpdParams& p, SM& sm, const float* Xs, float* g_w0,
                                                     float* g_b0, bool first, int tid, const float* Cx = nullptr) {
  if constexpr (ADAP) {
    // context rows of the first-layer gradient (AdapPolicy: dense inputs behind the features):
    // gW0[F + cc][j] = sum_b fma(dz1[j][b], ctx[cc][b], .), b ascending
    if (tid < p.C * HID) {
      const int cc = tid >> 6, j = tid & 63;
      float acc = 0.f;
      for (int b = 0; b < BT; ++b) acc = fmaf(BOX ? sm.H2[j * LDA + b] : sm.H2[b * LDT + j], Cx[cc * LDA + b], acc);
      acc_store(g_w0 + (p.sp.F + cc) * HID + j, acc, first);
    }
  }
  if constexpr (BOX) {
    row_sums(sm.H2, HID, g_b0, first, tid, UNT - HID);
    wgrad_first_box(sm.H2, Xs, p.sp.F, g_w0, first, tid);
    return;
  }
  constexpr int SLOTS_PER_WARP = (SM::OW + NWS - 1) / NWS;
  float4 csum[SLOTS_PER_WARP];
  if (tid >= UNT - HID) {  // warps NWS, NWS + 1: the bias chains, next to the other warps' walks
    const int j = tid - (UNT - HID);
    float s = 0.f;
    for (int b0 = 0; b0 < BT; b0 += 8) {  // 8 loads in flight, then the 8 adds of the chain
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = sm.H2[(b0 + u) * LDT + j];
#pragma unroll
      for (int u = 0; u < 8; ++u) s = s + x[u];
    }
    acc_store(g_b0 + j, s, first);
    sm.bc[j] = s;  // bc is free during the tower's backward pass
  }
  segsum_w1_walk(p, sm, sm.H2, g_w0, first, tid, csum);
  __syncthreads();  // bias chains complete
  segsum_w1_mode(p, sm, sm.bc, g_w0, first, tid, csum);
}

template <bool BOX, bool ADAP = false, class SM>
__device__ __forceinline__ void tower_backward(const UpdParams& p, SM& sm, const float* Xs,
                                               const float* Ha, const float* w1_s, float* g_w0,
                                               float* g_b0, float* g_w1, float* g_b1, int nb,
                                               bool first, int tid, long long& prof_last, int c,
                                               int pbase, const float* Cx = nullptr) {
  (void)nb;
  __syncthreads();  // D1 complete
  PTH_PROF(pbase + 0);
  if (tid < UNT / 2) {
    backprop64<!BOX, UNT / 2>(sm.D1, w1_s, Ha, sm.H2, tid);
  } else {
    wgrad64<UNT / 2>(sm.D1, Ha, g_w1, first, tid - UNT / 2);
    row_sums(sm.D1, HID, g_b1, first, tid, UNT - HID);
  }
  __syncthreads();  // dz1 complete (in H2)
  PTH_PROF(pbase + 1);  // backprop64 | wgrad64 + bias sums
  first_layer_backward<BOX, ADAP>(p, sm, Xs, g_w0, g_b0, first, tid, Cx);
}

template <bool BOX, bool WIDE = false, bool ADAP = false, bool MOD = false, bool MULT = false>
__global__ void __launch_bounds__(UNT) ppo_update_kernel(const __grid_constant__ UpdParams p) {
  static_assert(!(BOX && WIDE), "wide rows are a one-hot notion");
  static_assert(!((ADAP || MOD) && WIDE) && !(ADAP && MOD), "ADAP / Modular: one-hot rows of 32 slots or Box rows");
  static_assert(!MULT || ADAP, "AdapPolicyMult is an ADAP policy");
  constexpr bool PLAIN = WIDE || ADAP || MOD;
  using UpdSmem = UpdSmemT<WIDE, PLAIN>;
  constexpr int OW = UpdSmem::OW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  UpdSmem& sm = *reinterpret_cast<UpdSmem*>(smem_raw);
  // one more [64][LDA] tile behind the struct.  BOX: the observation rows, feature major (Xs).
  // One-hot: the value tower's first-layer activations (V1), computed next to the policy
  // tower's at the start of the tile and kept until the value tower runs.
  float* Xs = reinterpret_cast<float*>(smem_raw + sizeof(UpdSmem));
  float* V1 = BOX ? sm.H1 : Xs;
  // ADAP: behind Xs / V1, the context values of the tile's samples, [MAX_CTX][LDA]
  [[maybe_unused]] float* Cx = Xs + HID * LDA;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  const int c = blockIdx.x;
  const int P = p.P;
  const int64_t n_mb = (p.M + p.BS - 1) / p.BS;
  const int PS = (P + 3) & ~3;  // per-CTA stride of the partial sums: keeps every row 16-byte aligned
  float* part = p.part + (size_t)c * PS;
  const uint8_t* obs_s = reinterpret_cast<const uint8_t*>(sm.obs);
  if (tid < OW) sm.dmode[tid] = 0;  // slots beyond obs_len compare equal (observation rows are zero padded)

  // ------------------------------------------------ prologue: advantage statistics
  // One whole CTA per minibatch id (the lane order of the contract does not depend on WHICH CTA).
  // Multi-GPU: every rank holds the same gathered stream, so rank r computes the ids congruent to
  // r (mod world) only — 1 / world of the gathers (three dependent DRAM accesses per sample once
  // the stream outgrows L2) — and stores {mean, std} into every rank's exchange array as
  // {value, launch tag} pairs; readers poll the pair at the head of the minibatch.
  const uint32_t adv_tag = p.flag_epoch + 1u;
  for (int64_t id = p.rank + (int64_t)p.world * c; id < (int64_t)p.n_epochs * n_mb; id += (int64_t)p.world * G) {
    const int e = (int)(id / n_mb);
    const int64_t m = id % n_mb;
    const int64_t i0 = m * p.BS;
    const int64_t B = (i0 + p.BS <= p.M) ? p.BS : (p.M - i0);
    float mean = 0.f, stdv = 0.f;
    if (p.normalize && B > 1) {
      // 128 strided lanes (reduction contract), threads [BT, UNT) idle here
      // lane t < 128 adds elements t, t + 128, ... in that order (reduction contract).  The gathers
      // (permutation -> index -> advantage: three dependent DRAM accesses at large batches) are done by
      // ALL threads, 16 in flight each, into a shared-memory stage of CH elements; the lanes then add
      // their elements of the stage in order.
      constexpr int CH = 8192;  // floats staged per round (fits sm.H1)
      float* stage = sm.H1;
      float s = 0.f, q = 0.f;
      for (int pass = 0; pass < 2; ++pass) {
        for (int64_t base = 0; base < B; base += CH) {
          float v[CH / UNT];
#pragma unroll
          for (int u = 0; u < CH / UNT; ++u) {
            const int64_t k = base + tid + (int64_t)u * UNT;
            v[u] = k < B ? *reinterpret_cast<const float*>(p.adv + sample_offset(p, e, i0 + k) * p.f_stride) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < CH / UNT; ++u) stage[tid + u * UNT] = v[u];
          __syncthreads();
          if (tid < BT) {
            const int n = (int)((B - base < CH) ? (B - base) : CH);
            if (pass == 0) {
              for (int j = tid; j < n; j += BT) s = s + stage[j];
            } else {
              for (int j = tid; j < n; j += BT) {
                const float d = stage[j] - mean;
                q = fmaf(d, d, q);
              }
            }
          }
          __syncthreads();
        }
        if (pass == 0) mean = block_tree(s, sm.red, tid) / (float)B;
      }
      stdv = sqrtf(block_tree(q, sm.red, tid) / (float)(B - 1));
    }
    if (p.world == 1) {
      if (tid == 0) {
        p.advstat[2 * id] = mean;
        p.advstat[2 * id + 1] = stdv;
      }
    } else if (tid < 2 * p.world) {
      ll_store(p.advx[tid >> 1] + 2 * id + (tid & 1), (tid & 1) ? stdv : mean, adv_tag);
    }
  }
  grid.sync();

  long long prof_last = clock64();
  double b1pow = p.b1pow0, b2pow = p.b2pow0;
  [[maybe_unused]] double b1pow_v = p.b1pow0_v, b2pow_v = p.b2pow0_v;
  const float omb1 = (float)(1.0 - (double)p.b1), omb2 = (float)(1.0 - (double)p.b2);
  const float clip_lo = 1.0f - p.clip, clip_hi = 1.0f + p.clip;
  const int S = ((P + G - 1) / G + 3) / 4 * 4;

  for (int e = 0; e < p.n_epochs; ++e) {
    for (int64_t m = 0; m < n_mb; ++m) {
      const int64_t id = (int64_t)e * n_mb + m;
      const int64_t i0 = m * p.BS;
      const int64_t B = (i0 + p.BS <= p.M) ? p.BS : (p.M - i0);
      const float Bf = (float)B, invB = 1.0f / Bf;
      const bool norm = p.normalize && B > 1;
      const int W = p.world;
      float mean, stdv;
      if (W == 1) {
        mean = __ldcg(p.advstat + 2 * id);
        stdv = __ldcg(p.advstat + 2 * id + 1);
      } else {
        // computed by rank id mod W (prologue); every thread polls the two pairs (same address: one
        // L2 transaction per warp)
        const uint2* ax = p.advx[p.rank] + 2 * id;
        const uint2 a0 = ll_peek(ax), a1 = ll_peek(ax + 1);  // both in flight
        mean = ll_wait(ax, adv_tag, a0);
        stdv = ll_wait(ax + 1, adv_tag, a1);
      }
      const int64_t n_main = (B + BT - 1) / BT;
      // ADAP context loss (adap/util.py:97-131): K sampled contexts on S_eff sampled states of the
      // minibatch, as extra tiles behind the minibatch's own: a context tile holds SC = 128 / K
      // states x K contexts (column = k * ns + state), so every pair of distributions of a state
      // meets inside one tile
      [[maybe_unused]] int S_eff = 0, SC = 1, n_ctx_tiles = 0;
      if constexpr (ADAP) {
        if (p.ctx_coeff != 0.f && p.K >= 2) {
          S_eff = (int)(p.NS < B ? p.NS : B);
          SC = BT / p.K;
          n_ctx_tiles = (S_eff + SC - 1) / SC;
        }
      }
      const int64_t n_tiles = n_main + n_ctx_tiles;
      const int64_t local_tiles = (n_tiles - p.rank + W - 1) / W;  // tile t belongs to rank t mod W
      const int A = (int)(local_tiles < G ? local_tiles : G);

      PTH_PROF(0);  // loop head (prologue on the first pass)
      __syncthreads();
      // (issuing these loads before the first tile's sample gather and committing them behind it was measured:
      // slower — 51.8 vs 50.1 ms — the 24 registers held across the gather spill at the 128-register cap)
      static_assert(UNT == 512, "load_policy_512");
      load_policy_512<true>(sm.pol, p.params, p.lo, p.sp.L, tid);
      PTH_PROF(1);  // weights -> smem
      float cta_stat[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      [[maybe_unused]] float cta_ctx = 0.f;  // ADAP (thread 0): sum over this CTA's context tiles of sum_states sum_pairs exp(-KL)
      bool first = true;

      for (int64_t tau = p.rank + (int64_t)W * c; tau < n_tiles; tau += (int64_t)W * G) {
        const int64_t t0 = tau * BT;
        const bool ctile = ADAP && tau >= n_main;
        [[maybe_unused]] const int ct = (int)(tau - n_main);
        [[maybe_unused]] const int ns = ctile ? ((S_eff - ct * SC < SC) ? (S_eff - ct * SC) : SC) : 1;
        const int nb = ctile ? p.K * ns : (int)((B - t0 < BT) ? (B - t0) : BT);
        const bool valid = tid < nb;
        // flat offset of the sample in column `col` of this tile
        auto tile_off = [&](int col) -> int64_t {
          if constexpr (ADAP) {
            if (ctile) return sample_offset(p, e, i0 + __ldg(p.ctx_sidx + id * p.NS + ct * SC + col % ns));
          }
          return sample_offset(p, e, i0 + t0 + col);
        };
        // ---- gather this thread's sample
        uint32_t act = 0;
        float adv = 0.f, oldlp = 0.f, ret = 0.f;
        constexpr int OQ = OW / 16;  // uint4 per observation row
        uint4 orow[OQ];
#pragma unroll
        for (int i = 0; i < OQ; ++i) orow[i] = make_uint4(0, 0, 0, 0);
        constexpr int XQ = 16 / (UNT / BT);  // BOX: float4 of the sample's 64-float row per thread
        float4 xrow[XQ];                     // (threads b, b + 128, ... share sample b)
        if constexpr (BOX) {
          const int b = tid & (BT - 1);
#pragma unroll
          for (int i = 0; i < XQ; ++i) xrow[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (b < nb) {
            const int64_t offb = tile_off(b);
            const float4* q = reinterpret_cast<const float4*>(p.obs + offb * p.obs_stride) + (tid >> 7) * XQ;
#pragma unroll
            for (int i = 0; i < XQ; ++i) xrow[i] = __ldg(q + i);
          }
        }
        const int sb = tid & (BT - 1);  // the sample this thread works on in the per-sample phases
        const bool svalid = sb < nb;
        [[maybe_unused]] float cxv[MAX_CTX];
        if constexpr (ADAP) {
#pragma unroll
          for (int cc = 0; cc < MAX_CTX; ++cc) cxv[cc] = 0.f;
        }
        if (svalid) {
          const int64_t off = tile_off(sb);
          if constexpr (ADAP) {
            if (valid) {  // the context stored with the sample, or the sampled context k = column / ns
              const float* src = ctile ? p.ctx_draws + ((size_t)id * p.K + tid / ns) * p.C : p.ctx + off * p.C;
#pragma unroll
              for (int cc = 0; cc < MAX_CTX; ++cc)
                if (cc < p.C) cxv[cc] = __ldg(src + cc);
            }
          }
          if constexpr (!BOX) {
            if (valid) {
              const uint4* q = reinterpret_cast<const uint4*>(p.obs + off * p.obs_stride);
#pragma unroll
              for (int i = 0; i < OQ; ++i) orow[i] = __ldg(q + i);
            }
          }
          act = *reinterpret_cast<const uint32_t*>(p.actions + off * p.act_stride);
          adv = *reinterpret_cast<const float*>(p.adv + off * p.f_stride);
          oldlp = *reinterpret_cast<const float*>(p.old_logp + off * p.f_stride);
          if (valid) ret = *reinterpret_cast<const float*>(p.ret + off * p.f_stride);
        }
        const bool lane = tid < BT;  // threads [0, BT) own one sample each
        __syncthreads();  // previous tile done with sm.obs / Xs / Lg / H2
        if constexpr (BOX) {
          const int b = tid & (BT - 1), k0 = (tid >> 7) * (4 * XQ);
#pragma unroll
          for (int i = 0; i < XQ; ++i) {
            const float v[4] = {xrow[i].x, xrow[i].y, xrow[i].z, xrow[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = k0 + 4 * i + j;
              Xs[k * LDA + b] = k < p.sp.F ? v[j] : 0.f;
            }
          }
        } else if (lane) {
#pragma unroll
          for (int i = 0; i < OQ; ++i) *reinterpret_cast<uint4*>(&sm.obs[tid * (OW / 4) + 4 * i]) = orow[i];
          if (tid == 0) sm.n_staged = 0;
        }
        if constexpr (ADAP) {
          if (lane) {
#pragma unroll
            for (int cc = 0; cc < MAX_CTX; ++cc) Cx[cc * LDA + tid] = cxv[cc];
          }
        }
        __syncthreads();
        PTH_PROF(2);  // sample gather
        // (moving the sort into the shadow of the head phase, onto the warps without a head, was
        // measured: no gain — it competes with the head warps for issue slots)
        if constexpr (!BOX) {
          sort_slots(p, sm, obs_s, nb, tid & 31, tid >> 5, UNT / 32);
          __syncthreads();  // order / rcount / dmode / rowmap complete; the sort's scratch (rowpos) is free
          PTH_PROF(22);  // slot sort + row numbering
          if constexpr (!PLAIN) {
            chain_setup(p, sm, sm.H2, nb, tid);
            __syncthreads();
          }
          PTH_PROF(23);  // jb, stage copy, row positions | chain beginnings
        }

        // ================= ModularPolicy: forward of the main network and of every partner module
        [[maybe_unused]] ModScratch gsr;
        [[maybe_unused]] float mod_value = 0.f;  // lane threads: main value + the trained partner's value
        if constexpr (MOD) {
          const ModBlock mbk = mod_block(p.sp.L);
          gsr = mod_scratch(p.mod_ws + (size_t)c * mod_scratch_floats(p.Pn), p.Pn);
          const int L = p.sp.L;
          // main first layers -> a1p, a1v
          if constexpr (BOX) {
            first_layer_box<true, UNT>(p.sp.F, Xs, p.params + p.lo.w_pi0, sm.pol.b_pi0, gsr.a1p, tid);
            first_layer_box<true, UNT>(p.sp.F, Xs, p.params + p.lo.w_vf0, sm.pol.b_vf0, gsr.a1v, tid);
          } else if (tid < UNT / 2) {
            first_layer_onehot<true, UNT / 2, BT, 3, OW>(p.sp, obs_s, p.params + p.lo.w_pi0, sm.pol.b_pi0, gsr.a1p, tid);
          } else {
            first_layer_onehot<true, UNT / 2, BT, 3, OW>(p.sp, obs_s, p.params + p.lo.w_vf0, sm.pol.b_vf0, gsr.a1v,
                                                          tid - UNT / 2);
          }
          __syncthreads();
          // the previous tile left a partner's weights in the slots: the main matrices again
          load_policy<true>(sm.pol, p.params, p.lo, L, tid, UNT);
          __syncthreads();
          dense64<true, UNT, BT, LDA>(gsr.a1p, sm.pol.w_pi1, sm.pol.b_pi1, gsr.h2, tid);
          dense64<true, UNT, BT, LDA>(gsr.a1v, sm.pol.w_vf1, sm.pol.b_vf1, gsr.v2, tid);
          __syncthreads();
          logits_tile(gsr.h2, sm.pol, L, gsr.lgm, tid);
          if (tid < BT) mod_value = value_head(gsr.v2, sm.pol, tid);
          __syncthreads();
          for (int pq = 0; pq < p.Pn; ++pq) {
            const float* pb = p.params + p.lo.total + (size_t)pq * mbk.total;
            stage_rows64(sm.pol.w_pi1, pb + mbk.w_pi0, HID, tid);
            stage_rows64(sm.pol.w_vf1, pb + mbk.w_pi1, HID, tid);
            stage_rows64(sm.pol.w_act, pb + mbk.w_act, L, tid);
            stage_vec(sm.pol.b_pi1, pb + mbk.b_pi0, HID, tid);
            stage_vec(sm.pol.b_vf1, pb + mbk.b_pi1, HID, tid);
            stage_vec(sm.pol.b_act, pb + mbk.b_act, L, tid);
            __syncthreads();
            dense64<true, UNT, BT, LDA>(gsr.h2, sm.pol.w_pi1, sm.pol.b_pi1, gsr.q1 + (size_t)pq * TILE_F, tid);
            __syncthreads();
            dense64<true, UNT, BT, LDA>(gsr.q1 + (size_t)pq * TILE_F, sm.pol.w_vf1, sm.pol.b_vf1,
                                        gsr.q2 + (size_t)pq * TILE_F, tid);
            __syncthreads();
            logits_tile(gsr.q2 + (size_t)pq * TILE_F, sm.pol, L, gsr.lgp + (size_t)pq * LTILE_F, tid);
            __syncthreads();
          }
          {  // value branch of the trained partner (it reads latent_pi too: modular/policies.py:279, 375)
            const float* pb = p.params + p.lo.total + (size_t)p.p0 * mbk.total;
            stage_rows64(sm.pol.w_pi1, pb + mbk.w_vf0, HID, tid);
            stage_rows64(sm.pol.w_vf1, pb + mbk.w_vf1, HID, tid);
            stage_vec(sm.pol.b_pi1, pb + mbk.b_vf0, HID, tid);
            stage_vec(sm.pol.b_vf1, pb + mbk.b_vf1, HID, tid);
            stage_vec(sm.pol.w_val, pb + mbk.w_val, HID, tid);
            if (tid == 0) sm.pol.b_val = __ldcg(pb + mbk.b_val);
            __syncthreads();
            dense64<true, UNT, BT, LDA>(gsr.h2, sm.pol.w_pi1, sm.pol.b_pi1, gsr.r1, tid);
            __syncthreads();
            dense64<true, UNT, BT, LDA>(gsr.r1, sm.pol.w_vf1, sm.pol.b_vf1, gsr.r2, tid);
            __syncthreads();
            if (tid < BT) mod_value = mod_value + value_head(gsr.r2, sm.pol, tid);
          }
          // composed logits of the trained partner: what PPO's losses see
          for (int i = tid; i < L * BT; i += UNT) {
            const int l = i >> 7, b = i & (BT - 1);
            sm.Lg[l * LDA + b] = gsr.lgm[l * LDA + b] + gsr.lgp[(size_t)p.p0 * LTILE_F + l * LDA + b];
          }
        }
        // ================= AdapPolicyMult: forward of both towers (a context tile: the policy tower only)
        [[maybe_unused]] MultScratch msr;
        if constexpr (MULT) {
          const int C = p.C, L = p.sp.L;
          msr = mult_scratch(p.mod_ws + (size_t)c * mult_scratch_floats(C), C);
          float* slotW = sm.H1;               // [64][LDW] staging slot of a scaling sub-layer's weights
          float* slotB = sm.H1 + HID * LDW;   // its 64 biases
          if constexpr (BOX) {
            first_layer_box<true, UNT>(p.sp.F, Xs, p.params + p.lo.w_pi0, sm.pol.b_pi0, msr.x[0], tid);
            first_layer_box<true, UNT>(p.sp.F, Xs, p.params + p.lo.w_vf0, sm.pol.b_vf0, msr.x[1], tid);
          } else if (tid < UNT / 2) {
            first_layer_onehot<true, UNT / 2, BT, 3, OW>(p.sp, obs_s, p.params + p.lo.w_pi0, sm.pol.b_pi0, msr.x[0], tid);
          } else {
            first_layer_onehot<true, UNT / 2, BT, 3, OW>(p.sp, obs_s, p.params + p.lo.w_vf0, sm.pol.b_vf0, msr.x[1],
                                                          tid - UNT / 2);
          }
          __syncthreads();
          for (int t = 0; t < (ctile ? 1 : 2); ++t) {
            for (int cc = 0; cc < C; ++cc) {  // s_c = tanh(Ws_c x + bs_c): rows j C + cc of the scaling layer
              stage_rows64_strided(slotW, p.params + p.mult_w[t] + cc * HID, C, tid);
              stage_vec_strided(slotB, p.params + p.mult_b[t] + cc, HID, C, tid);
              __syncthreads();
              dense64<true, UNT, BT, LDA>(msr.x[t], slotW, slotB, msr.s[t] + (size_t)cc * TILE_F, tid);
              __syncthreads();
            }
            for (int i = tid; i < HID * BT; i += UNT) {  // y = x + sum_c s_c ctx_c
              const int k = i >> 7, b = i & (BT - 1);
              float acc = msr.x[t][k * LDA + b];
              for (int cc = 0; cc < C; ++cc) acc = fmaf(msr.s[t][(size_t)cc * TILE_F + k * LDA + b], Cx[cc * LDA + b], acc);
              msr.y[t][k * LDA + b] = acc;
            }
            __syncthreads();
            dense64<true, UNT, BT, LDA>(msr.y[t], t ? sm.pol.w_vf1 : sm.pol.w_pi1, t ? sm.pol.b_vf1 : sm.pol.b_pi1,
                                        msr.h2[t], tid);
            __syncthreads();
          }
          logits_tile(msr.h2[0], sm.pol, L, sm.Lg, tid);
          if (!ctile && tid < BT) mod_value = value_head(msr.h2[1], sm.pol, tid);
        }
        // ================= policy tower: forward
        if constexpr (MOD || MULT) {
        } else if constexpr (BOX) {
          first_layer_box<true, UNT, BT, !ADAP>(p.sp.F, Xs, p.params + p.lo.w_pi0, sm.pol.b_pi0, sm.H1, tid);
        } else {
          // both towers' first layers (latency-bound row gathers) side by side, one per CTA half
          if constexpr (PLAIN) {
            if (tid < UNT / 2)
              first_layer_onehot<true, UNT / 2, BT, 3, OW>(p.sp, obs_s, p.params + p.lo.w_pi0, sm.pol.b_pi0, sm.H1,
                                                            tid, !ADAP);
            else
              first_layer_onehot<true, UNT / 2, BT, 3, OW>(p.sp, obs_s, p.params + p.lo.w_vf0, sm.pol.b_vf0, V1,
                                                            tid - UNT / 2, !ADAP);
          } else if (tid < UNT / 2) {
            first_layer_chain<UNT / 2>(p, sm, obs_s, p.params + p.lo.w_pi0, sm.H2, 0, sm.pchain[0], sm.H1, tid);
          } else {
            first_layer_chain<UNT / 2>(p, sm, obs_s, p.params + p.lo.w_vf0, sm.H2, 1, sm.pchain[1], V1,
                                       tid - UNT / 2);
          }
        }
        __syncthreads();
        if constexpr (ADAP && !MULT) {
          // the context inputs continue the chains, then tanh (one-hot: one tower per CTA half)
          if constexpr (BOX)
            context_columns_tanh<true>(p.C, Cx, p.params + p.lo.w_pi0 + p.sp.F * HID, sm.H1, tid >> 5, UNT / 32,
                                       tid & 31);
          else
            context_columns_tanh<true>(p.C, Cx, p.params + (tid < UNT / 2 ? p.lo.w_pi0 : p.lo.w_vf0) + p.sp.F * HID,
                                       tid < UNT / 2 ? sm.H1 : V1, (tid >> 5) & (UNT / 64 - 1), UNT / 64, tid & 31);
          __syncthreads();
        }
        PTH_PROF(3);  // pi first layer (one-hot: both towers' first layers)
        if constexpr (!(MOD || MULT))
          dense64<true, UNT / 2, BT / 2, LDA>(sm.H1 + (tid >> 8) * (BT / 2), sm.pol.w_pi1, sm.pol.b_pi1,
                                               sm.H2 + (tid >> 8) * (BT / 2), tid & (UNT / 2 - 1));
        __syncthreads();
        PTH_PROF(4);  // pi hidden layer
        float s_pl = 0.f, s_e = 0.f, s_kl = 0.f, s_cf = 0.f, s_v = 0.f;
        if constexpr (!(MOD || MULT)) logits_tile(sm.H2, sm.pol, p.sp.L, sm.Lg, tid);
        __syncthreads();
        if (!ctile) {
          // ---- per-sample losses and d loss / d logits.  Thread group h (threads [128h, 128h + 128))
          // evaluates head h of sample sb: same values as dist_eval (pth_mlp.cuh), with every exp
          // computed once (the softmax probabilities are parked in the sample's column of sm.D1,
          // free until dz2 is written).  The heads' log-probs / entropies meet in rows 32.. of D1
          // and every group adds them in head order, as the one-thread-per-sample version did.
          static_assert(UNT / BT >= PTH_MAX_HEADS, "one thread group per head");
          float* Pr = sm.D1;
          float* hs = sm.D1 + MAXL * LDA;  // [2 * PTH_MAX_HEADS][LDA] scratch
          const int hh = tid >> 7;
          const bool mine = hh < p.sp.n_heads;
          int off = 0, n = 0, a_h = 0;
          float mx = 0.f, logS = 0.f, ent = 0.f;
          if (mine) {
            for (int h = 0; h < hh; ++h) off += p.sp.head_n[h];
            n = p.sp.head_n[hh];
            a_h = (int)((act >> (8 * hh)) & 0xffu);
            mx = sm.Lg[off * LDA + sb];
            for (int i = 1; i < n; ++i) {
              const float z = sm.Lg[(off + i) * LDA + sb];
              mx = z > mx ? z : mx;
            }
            float Ssum = 0.f;
            for (int i = 0; i < n; ++i) {
              const float ex = pth_expf(sm.Lg[(off + i) * LDA + sb] - mx);
              Pr[(off + i) * LDA + sb] = ex;
              Ssum = Ssum + ex;
            }
            logS = pth_logf(Ssum);
            for (int i = 0; i < n; ++i) {
              const float lp = (sm.Lg[(off + i) * LDA + sb] - mx) - logS;
              const float pi = Pr[(off + i) * LDA + sb] / Ssum;
              Pr[(off + i) * LDA + sb] = pi;
              ent = fmaf(-pi, lp, ent);
            }
            hs[hh * LDA + sb] = (sm.Lg[(off + a_h) * LDA + sb] - mx) - logS;
            hs[(PTH_MAX_HEADS + hh) * LDA + sb] = ent;
          }
          __syncthreads();
          float logp = 0.f, entropy = 0.f;
#pragma unroll
          for (int h = 0; h < PTH_MAX_HEADS; ++h) {
            if (h < p.sp.n_heads) {
              logp = logp + hs[h * LDA + sb];
              entropy = entropy + hs[(PTH_MAX_HEADS + h) * LDA + sb];
            }
          }
          float glp;
          const float gH = svalid ? -(p.ent_coef * invB) : 0.f;
          if (p.loss_kind == PTH_LOSS_BC) {
            // bc.py:297-315: loss = -mean(log_prob) - ent_weight * mean(entropy) (+ l2, in Adam)
            glp = svalid ? -invB : 0.f;
            if (lane) {
              s_pl = valid ? logp : 0.f;            // column 0: neglogp
              s_e = valid ? entropy : 0.f;
              s_kl = valid ? pth_expf(logp) : 0.f;  // column 3: prob_true_act
              s_cf = 0.f;
            }
          } else {
            if (norm) adv = (adv - mean) / (stdv + 1e-8f);
            const float lr_ = logp - oldlp;
            const float ratio = pth_expf(lr_);
            const float pl1 = adv * ratio;
            const float rc = fminf(fmaxf(ratio, clip_lo), clip_hi);
            const float pl2 = adv * rc;
            const bool inside = ratio >= clip_lo && ratio <= clip_hi;
            const bool gmask = inside || (pl1 < pl2);
            glp = (svalid && gmask) ? -((adv * ratio) * invB) : 0.f;
            if (lane) {
              s_pl = valid ? fminf(pl1, pl2) : 0.f;
              s_e = valid ? entropy : 0.f;
              s_kl = valid ? (ratio - 1.0f) - lr_ : 0.f;
              s_cf = (valid && fabsf(ratio - 1.0f) > p.clip) ? 1.f : 0.f;
            }
          }
          if (mine) {
            for (int i = 0; i < n; ++i) {
              const float zi = sm.Lg[(off + i) * LDA + sb];
              const float lp = (zi - mx) - logS;
              const float pi = Pr[(off + i) * LDA + sb];
              const float t1 = (i == a_h ? 1.0f : 0.0f) - pi;
              const float dzv = glp * t1;
              const float t2 = (gH * pi) * (lp + ent);
              sm.Lg[(off + i) * LDA + sb] = dzv - t2;
            }
          }
        }
        if constexpr (ADAP) {
          if (ctile) {
            // ---- context loss of this tile's states (adap/util.py:97-131): mean over the K (K - 1) / 2 pairs
            // of mean_states exp(-KL(dist_a || dist_b)), dist_k = the policy head under context k.
            // Thread group h turns head h of every column into probabilities (D1 rows [0, L)) and
            // log-probabilities (D1 rows [32, 32 + L)) and clears the column's dlogits; then one thread
            // per state walks the pairs in ascending order, adding into the two columns' dlogits
            //   d KL / d z_a = p_a ((lp_a - lp_b) - KL_h),   d KL / d z_b = p_b - p_a   (per head h).
            float* Pr = sm.D1;
            float* LP = sm.D1 + MAXL * LDA;
            static_assert(2 * MAXL <= HID, "probabilities and log-probabilities share D1");
            const int hh = tid >> 7;
            if (hh < p.sp.n_heads) {
              int off = 0;
              for (int h = 0; h < hh; ++h) off += p.sp.head_n[h];
              const int n = p.sp.head_n[hh];
              float mx = sm.Lg[off * LDA + sb];
              for (int i = 1; i < n; ++i) {
                const float z = sm.Lg[(off + i) * LDA + sb];
                mx = z > mx ? z : mx;
              }
              float Ssum = 0.f;
              for (int i = 0; i < n; ++i) {
                const float ex = pth_expf(sm.Lg[(off + i) * LDA + sb] - mx);
                Pr[(off + i) * LDA + sb] = ex;
                Ssum = Ssum + ex;
              }
              const float logS = pth_logf(Ssum);
              for (int i = 0; i < n; ++i) {
                LP[(off + i) * LDA + sb] = (sm.Lg[(off + i) * LDA + sb] - mx) - logS;
                Pr[(off + i) * LDA + sb] = Pr[(off + i) * LDA + sb] / Ssum;
                sm.Lg[(off + i) * LDA + sb] = 0.f;
              }
            }
            __syncthreads();
            if (tid < ns) {
              const float NPf = (float)(p.K * (p.K - 1) / 2), Sf = (float)S_eff;
              float esum = 0.f;
              for (int ka = 0; ka < p.K; ++ka)
                for (int kb = ka + 1; kb < p.K; ++kb) {
                  const int ca = ka * ns + tid, cb = kb * ns + tid;
                  float klh[PTH_MAX_HEADS], kl = 0.f;
                  int o = 0;
#pragma unroll
                  for (int h = 0; h < PTH_MAX_HEADS; ++h) {
                    klh[h] = 0.f;
                    if (h < p.sp.n_heads) {
                      float acc = 0.f;
                      for (int i = 0; i < p.sp.head_n[h]; ++i)
                        acc = fmaf(Pr[(o + i) * LDA + ca], LP[(o + i) * LDA + ca] - LP[(o + i) * LDA + cb], acc);
                      klh[h] = acc;
                      kl = kl + acc;
                      o += p.sp.head_n[h];
                    }
                  }
                  const float ex = pth_expf(-kl);
                  esum = esum + ex;
                  const float g = -(((p.ctx_coeff * ex) / NPf) / Sf);
                  o = 0;
#pragma unroll
                  for (int h = 0; h < PTH_MAX_HEADS; ++h) {
                    if (h < p.sp.n_heads) {
                      for (int i = 0; i < p.sp.head_n[h]; ++i) {
                        const float pa = Pr[(o + i) * LDA + ca], pb = Pr[(o + i) * LDA + cb];
                        const float da = pa * ((LP[(o + i) * LDA + ca] - LP[(o + i) * LDA + cb]) - klh[h]);
                        const float db = pb - pa;
                        sm.Lg[(o + i) * LDA + ca] = sm.Lg[(o + i) * LDA + ca] + g * da;
                        sm.Lg[(o + i) * LDA + cb] = sm.Lg[(o + i) * LDA + cb] + g * db;
                      }
                      o += p.sp.head_n[h];
                    }
                  }
                }
              sm.bc[tid] = esum;
            }
            __syncthreads();
            if (tid == 0) {  // the tile's states in ascending order, then this CTA's context tiles in order
              float t = 0.f;
              for (int sl = 0; sl < ns; ++sl) t = t + sm.bc[sl];
              cta_ctx = first ? t : cta_ctx + t;
            }
          }
        }
        __syncthreads();  // dlogits complete
        PTH_PROF(5);  // action head + losses + dlogits
        if constexpr (MULT) {
          // ================= AdapPolicyMult: backward (contract: oracle/pth_oracle_update.inc mult_tower_backward).
          // sm.Lg holds d loss / d logits (PPO's, or the context loss's on a context tile).
          const int C = p.C, L = p.sp.L;
          float* slotW = sm.H1;
          float* DV = reinterpret_cast<float*>(sm.rowpos);  // [BT] d loss / d value
          for (int i = tid; i < TILE_F; i += UNT) msr.zero[i] = 0.f;
          if (!ctile && lane) {
            const float dret = ret - mod_value;
            s_v = valid ? dret * dret : 0.f;
            DV[tid] = valid ? ((p.vf_coef * 2.0f) * (mod_value - ret)) * invB : 0.f;
          }
          // one tower, given dz2 in sm.D1: second layer reads y; dy = W1^T dz2; scaling sub-layers c ascending:
          // dzs = (dy ctx_c)(1 - s_c^2), gWs rows j C + c, dx += Ws_c^T dzs; dz1 = dx (1 - x^2) -> first-layer gradients
          auto tower = [&](int t) {
            float* g_w1 = part + (t ? p.lo.w_vf1 : p.lo.w_pi1);
            float* g_b1 = part + (t ? p.lo.b_vf1 : p.lo.b_pi1);
            wgrad64<UNT>(sm.D1, msr.y[t], g_w1, first, tid);
            row_sums(sm.D1, HID, g_b1, first, tid, UNT - HID);
            backprop64<false, UNT>(sm.D1, t ? sm.pol.w_vf1 : sm.pol.w_pi1, msr.zero, msr.dy, tid);
            __syncthreads();
            for (int i = tid; i < HID * BT; i += UNT) {
              const int k = i >> 7, b = i & (BT - 1);
              msr.dx[k * LDA + b] = msr.dy[k * LDA + b];
            }
            for (int cc = 0; cc < C; ++cc) {
              stage_rows64_strided(slotW, p.params + p.mult_w[t] + cc * HID, C, tid);
              const float* sc = msr.s[t] + (size_t)cc * TILE_F;
              for (int i = tid; i < HID * BT; i += UNT) {
                const int k = i >> 7, b = i & (BT - 1);
                const float sv_ = sc[k * LDA + b];
                sm.D1[k * LDA + b] = (msr.dy[k * LDA + b] * Cx[cc * LDA + b]) * (1.0f - sv_ * sv_);
              }
              __syncthreads();
              wgrad64<UNT>(sm.D1, msr.x[t], part + p.mult_w[t] + cc * HID, first, tid, C * HID);
              row_sums(sm.D1, HID, part + p.mult_b[t] + cc, first, tid, UNT - HID, C);
              backprop64<false, UNT>(sm.D1, slotW, msr.zero, msr.tmp, tid);
              __syncthreads();
              for (int i = tid; i < HID * BT; i += UNT) {
                const int k = i >> 7, b = i & (BT - 1);
                msr.dx[k * LDA + b] = msr.dx[k * LDA + b] + msr.tmp[k * LDA + b];
              }
              __syncthreads();
            }
            for (int i = tid; i < HID * BT; i += UNT) {
              const int k = i >> 7, b = i & (BT - 1);
              const float xv = msr.x[t][k * LDA + b];
              sm.H2[BOX ? (k * LDA + b) : (b * LDT + k)] = msr.dx[k * LDA + b] * (1.0f - xv * xv);
            }
            __syncthreads();
            first_layer_backward<BOX, false>(p, sm, Xs, part + (t ? p.lo.w_vf0 : p.lo.w_pi0),
                                             part + (t ? p.lo.b_vf0 : p.lo.b_pi0), first, tid);
            __syncthreads();
          };
          __syncthreads();
          if (tid >= UNT / 2) {
            head_wgrad(sm.Lg, msr.h2[0], L, part + p.lo.w_act, first, tid - UNT / 2);
            row_sums(sm.Lg, L, part + p.lo.b_act, first, tid, UNT - MAXL);
          }
          head_backprop<true>(sm.Lg, sm.pol.w_act, msr.h2[0], L, sm.D1, tid);
          __syncthreads();
          tower(0);
          if (ctile) {
            if (first) {  // no value tower on a context tile: its sums start at zero
              for (int i = p.lo.w_vf0 + tid; i < p.lo.w_act; i += UNT) part[i] = 0.f;
              for (int i = p.lo.w_val + tid; i < P; i += UNT) part[i] = 0.f;
            }
            first = false;
            continue;
          }
          if (lane) {
            const float dv = DV[tid];
#pragma unroll 8
            for (int k = 0; k < HID; ++k) {
              const float h = msr.h2[1][k * LDA + tid];
              sm.D1[k * LDA + tid] = (sm.pol.w_val[k] * dv) * (1.0f - h * h);
            }
          } else if (tid < BT + HID) {
            const int k = tid - BT;
            float acc = 0.f;
            for (int b0 = 0; b0 < BT; b0 += 4) {
              const float4 d = *reinterpret_cast<const float4*>(DV + b0);
              const float4 h = *reinterpret_cast<const float4*>(msr.h2[1] + k * LDA + b0);
              acc = fmaf(d.x, h.x, acc);
              acc = fmaf(d.y, h.y, acc);
              acc = fmaf(d.z, h.z, acc);
              acc = fmaf(d.w, h.w, acc);
            }
            acc_store(part + p.lo.w_val + k, acc, first);
          } else if (tid == BT + HID) {
            float s_ = 0.f;
            for (int b = 0; b < BT; ++b) s_ = s_ + DV[b];
            acc_store(part + p.lo.b_val, s_, first);
          }
          __syncthreads();
          tower(1);
          {
            float sv[5] = {s_pl, s_v, s_e, s_kl, s_cf};
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) sv[i] = sv[i] + __shfl_xor_sync(0xffffffffu, sv[i], d);
            __syncthreads();
            if ((tid & 31) == 0 && tid < BT)
#pragma unroll
              for (int i = 0; i < 5; ++i) sm.bc[i * 4 + (tid >> 5)] = sv[i];
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const float ts = ((sm.bc[i * 4] + sm.bc[i * 4 + 1]) + sm.bc[i * 4 + 2]) + sm.bc[i * 4 + 3];
              cta_stat[i] = first ? ts : cta_stat[i] + ts;
            }
          }
          first = false;
          continue;
        }
        if constexpr (MOD) {
          // ================= ModularAlgorithm: the marginal regulariser and the backward pass through
          // every module (contract: oracle/pth_oracle_modular.inc).  sm.Lg holds d PPO-loss / d composed logits.
          const ModBlock mbk = mod_block(p.sp.L);
          const int L = p.sp.L, Pn = p.Pn;
          const float Pnf = (float)Pn;
          float* DV = reinterpret_cast<float*>(sm.rowpos);  // [BT] d loss / d value (the sort's scratch is free)
          float s_mg = 0.f;
          if (lane) {
            const float dret = ret - mod_value;
            s_v = valid ? dret * dret : 0.f;
            DV[tid] = valid ? ((p.vf_coef * 2.0f) * (mod_value - ret)) * invB : 0.f;
            // softmax over ALL logits jointly (modular/learn.py:311-312): main, and main + partner[p]
            auto softmax_col = [&](const float* a, const float* b2, float* out) {
              float mx = a[tid] + (b2 ? b2[tid] : 0.f);
              for (int l = 1; l < L; ++l) {
                const float x = b2 ? a[l * LDA + tid] + b2[l * LDA + tid] : a[l * LDA + tid];
                mx = x > mx ? x : mx;
              }
              float S = 0.f;
              for (int l = 0; l < L; ++l) {
                const float x = b2 ? a[l * LDA + tid] + b2[l * LDA + tid] : a[l * LDA + tid];
                const float ex = pth_expf(x - mx);
                out[l * LDA + tid] = ex;
                S = S + ex;
              }
              for (int l = 0; l < L; ++l) out[l * LDA + tid] = out[l * LDA + tid] / S;
            };
            {
              float mx = gsr.lgm[tid];
              for (int l = 1; l < L; ++l) mx = gsr.lgm[l * LDA + tid] > mx ? gsr.lgm[l * LDA + tid] : mx;
              float S = 0.f;
              for (int l = 0; l < L; ++l) {
                const float ex = pth_expf(gsr.lgm[l * LDA + tid] - mx);
                gsr.pm[l * LDA + tid] = ex;
                S = S + ex;
              }
              for (int l = 0; l < L; ++l) gsr.pm[l * LDA + tid] = gsr.pm[l * LDA + tid] / S;
            }
            for (int pq = 0; pq < Pn; ++pq) softmax_col(gsr.lgm, gsr.lgp + (size_t)pq * LTILE_F, gsr.pc + (size_t)pq * LTILE_F);
            const float gcoef = valid ? p.marg_coef * invB : 0.f;
            float* G_ = gsr.tmp;  // d regulariser / d (main marginal - composed marginal), row l
            float absum = 0.f, t = 0.f;
            for (int l = 0; l < L; ++l) {
              const float pml = gsr.pm[l * LDA + tid];
              float sm_ = pml, sc_ = gsr.pc[l * LDA + tid];
              for (int pq = 1; pq < Pn; ++pq) {
                sm_ = sm_ + pml;
                sc_ = sc_ + gsr.pc[(size_t)pq * LTILE_F + l * LDA + tid];
              }
              const float diff = sm_ / Pnf - sc_ / Pnf;
              absum = absum + fabsf(diff);
              const float g = diff > 0.f ? gcoef : (diff < 0.f ? -gcoef : 0.f);
              G_[l * LDA + tid] = g;
              t = fmaf(g, pml, t);
            }
            s_mg = valid ? absum : 0.f;
            for (int l = 0; l < L; ++l) {
              const float pml = gsr.pm[l * LDA + tid];
              gsr.dlm[l * LDA + tid] = sm.Lg[l * LDA + tid] + pml * (G_[l * LDA + tid] - t);
            }
            for (int pq = 0; pq < Pn; ++pq) {
              const float* pc = gsr.pc + (size_t)pq * LTILE_F;
              float* dl = gsr.dlp + (size_t)pq * LTILE_F;
              float tp = 0.f;
              for (int l = 0; l < L; ++l) tp = fmaf(-(G_[l * LDA + tid] / Pnf), pc[l * LDA + tid], tp);
              for (int l = 0; l < L; ++l) {
                const float dzp = pc[l * LDA + tid] * (-(G_[l * LDA + tid] / Pnf) - tp);
                gsr.dlm[l * LDA + tid] = gsr.dlm[l * LDA + tid] + dzp;
                dl[l * LDA + tid] = pq == p.p0 ? sm.Lg[l * LDA + tid] + dzp : dzp;
              }
            }
          }
          // the zero tile (backprop64 without an activation derivative multiplies by 1 - 0^2)
          for (int i = tid; i < TILE_F; i += UNT) gsr.zero[i] = 0.f;
          // ---- main action head, and its share of d loss / d latent_pi (accumulated in sm.H1)
          stage_rows64(sm.pol.w_act, p.params + p.lo.w_act, L, tid);
          __syncthreads();
          if (tid >= UNT / 2) {
            head_wgrad(gsr.dlm, gsr.h2, L, part + p.lo.w_act, first, tid - UNT / 2);
            row_sums(gsr.dlm, L, part + p.lo.b_act, first, tid, UNT - MAXL);
          }
          head_backprop<false>(gsr.dlm, sm.pol.w_act, nullptr, L, sm.H1, tid);
          __syncthreads();
          // one 64 x 64 tanh layer backward: weight / bias gradients from (Dz, input activations A), then
          // Out = (W^T Dz) * (1 - Act^2)  (Act = the zero tile: no activation derivative)
          auto layer_backward = [&](const float* Dz, const float* A, const float* Wsm, const float* Act, float* Out,
                                    float* gW, float* gB) {
            wgrad64<UNT>(Dz, A, gW, first, tid);
            row_sums(Dz, HID, gB, first, tid, UNT - HID);
            backprop64<false, UNT>(Dz, Wsm, Act, Out, tid);
            __syncthreads();
          };
          auto add_into_dh2 = [&]() {
            for (int i = tid; i < HID * BT; i += UNT) {
              const int k = i >> 7, b = i & (BT - 1);
              sm.H1[k * LDA + b] = sm.H1[k * LDA + b] + gsr.tmp[k * LDA + b];
            }
            __syncthreads();
          };
          // ---- partner policy branches, p ascending
          for (int pq = 0; pq < Pn; ++pq) {
            const float* pb = p.params + p.lo.total + (size_t)pq * mbk.total;
            float* gb = part + p.lo.total + (size_t)pq * mbk.total;
            const float* dl = gsr.dlp + (size_t)pq * LTILE_F;
            const float* q1 = gsr.q1 + (size_t)pq * TILE_F;
            const float* q2 = gsr.q2 + (size_t)pq * TILE_F;
            stage_rows64(sm.pol.w_act, pb + mbk.w_act, L, tid);
            stage_rows64(sm.pol.w_pi1, pb + mbk.w_pi0, HID, tid);
            stage_rows64(sm.pol.w_vf1, pb + mbk.w_pi1, HID, tid);
            __syncthreads();
            if (tid >= UNT / 2) {
              head_wgrad(dl, q2, L, gb + mbk.w_act, first, tid - UNT / 2);
              row_sums(dl, L, gb + mbk.b_act, first, tid, UNT - MAXL);
            }
            head_backprop<true>(dl, sm.pol.w_act, q2, L, sm.D1, tid);
            __syncthreads();
            layer_backward(sm.D1, q1, sm.pol.w_vf1, q1, sm.H2, gb + mbk.w_pi1, gb + mbk.b_pi1);
            layer_backward(sm.H2, gsr.h2, sm.pol.w_pi1, gsr.zero, gsr.tmp, gb + mbk.w_pi0, gb + mbk.b_pi0);
            add_into_dh2();
          }
          // ---- value branch of the trained partner
          {
            const float* pb = p.params + p.lo.total + (size_t)p.p0 * mbk.total;
            float* gb = part + p.lo.total + (size_t)p.p0 * mbk.total;
            stage_rows64(sm.pol.w_pi1, pb + mbk.w_vf0, HID, tid);
            stage_rows64(sm.pol.w_vf1, pb + mbk.w_vf1, HID, tid);
            stage_vec(sm.pol.w_val, pb + mbk.w_val, HID, tid);
            __syncthreads();
            if (lane) {
              const float dv = DV[tid];
#pragma unroll 8
              for (int k = 0; k < HID; ++k) {
                const float h = gsr.r2[k * LDA + tid];
                sm.D1[k * LDA + tid] = (sm.pol.w_val[k] * dv) * (1.0f - h * h);
              }
            } else if (tid < BT + HID) {
              const int k = tid - BT;
              float acc = 0.f;
              for (int b0 = 0; b0 < BT; b0 += 4) {
                const float4 d = *reinterpret_cast<const float4*>(DV + b0);
                const float4 h = *reinterpret_cast<const float4*>(gsr.r2 + k * LDA + b0);
                acc = fmaf(d.x, h.x, acc);
                acc = fmaf(d.y, h.y, acc);
                acc = fmaf(d.z, h.z, acc);
                acc = fmaf(d.w, h.w, acc);
              }
              acc_store(gb + mbk.w_val + k, acc, first);
            } else if (tid == BT + HID) {
              float s_ = 0.f;
              for (int b = 0; b < BT; ++b) s_ = s_ + DV[b];
              acc_store(gb + mbk.b_val, s_, first);
            }
            __syncthreads();
            layer_backward(sm.D1, gsr.r1, sm.pol.w_vf1, gsr.r1, sm.H2, gb + mbk.w_vf1, gb + mbk.b_vf1);
            layer_backward(sm.H2, gsr.h2, sm.pol.w_pi1, gsr.zero, gsr.tmp, gb + mbk.w_vf0, gb + mbk.b_vf0);
            add_into_dh2();
          }
          // ---- main towers (the main matrices back in their slots)
          load_policy<true>(sm.pol, p.params, p.lo, L, tid, UNT);
          for (int i = tid; i < HID * BT; i += UNT) {
            const int k = i >> 7, b = i & (BT - 1);
            const float h = gsr.h2[k * LDA + b];
            sm.D1[k * LDA + b] = sm.H1[k * LDA + b] * (1.0f - h * h);
          }
          tower_backward<BOX>(p, sm, Xs, gsr.a1p, sm.pol.w_pi1, part + p.lo.w_pi0, part + p.lo.b_pi0, part + p.lo.w_pi1,
                              part + p.lo.b_pi1, nb, first, tid, prof_last, c, 18);
          __syncthreads();
          if (lane) {
            const float dv = DV[tid];
#pragma unroll 8
            for (int k = 0; k < HID; ++k) {
              const float h = gsr.v2[k * LDA + tid];
              sm.D1[k * LDA + tid] = (sm.pol.w_val[k] * dv) * (1.0f - h * h);
            }
          } else if (tid < BT + HID) {
            const int k = tid - BT;
            float acc = 0.f;
            for (int b0 = 0; b0 < BT; b0 += 4) {
              const float4 d = *reinterpret_cast<const float4*>(DV + b0);
              const float4 h = *reinterpret_cast<const float4*>(gsr.v2 + k * LDA + b0);
              acc = fmaf(d.x, h.x, acc);
              acc = fmaf(d.y, h.y, acc);
              acc = fmaf(d.z, h.z, acc);
              acc = fmaf(d.w, h.w, acc);
            }
            acc_store(part + p.lo.w_val + k, acc, first);
          } else if (tid == BT + HID) {
            float s_ = 0.f;
            for (int b = 0; b < BT; ++b) s_ = s_ + DV[b];
            acc_store(part + p.lo.b_val, s_, first);
          }
          tower_backward<BOX>(p, sm, Xs, gsr.a1v, sm.pol.w_vf1, part + p.lo.w_vf0, part + p.lo.b_vf0, part + p.lo.w_vf1,
                              part + p.lo.b_vf1, nb, first, tid, prof_last, c, 20);
          // ---- tile statistics: six trees (the sixth: the marginal regulariser's per-sample sums)
          {
            float sv[6] = {s_pl, s_v, s_e, s_kl, s_cf, s_mg};
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) sv[i] = sv[i] + __shfl_xor_sync(0xffffffffu, sv[i], d);
            __syncthreads();
            if ((tid & 31) == 0 && tid < BT)
#pragma unroll
              for (int i = 0; i < 6; ++i) sm.bc[i * 4 + (tid >> 5)] = sv[i];
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              const float ts = ((sm.bc[i * 4] + sm.bc[i * 4 + 1]) + sm.bc[i * 4 + 2]) + sm.bc[i * 4 + 3];
              if (i < 5)
                cta_stat[i] = first ? ts : cta_stat[i] + ts;
              else
                cta_ctx = first ? ts : cta_ctx + ts;
            }
          }
          first = false;
          continue;
        }
        // ================= policy tower: backward
        // upper half: head weight / bias gradients; lower half: dz2, two threads per sample
        // (32 hidden units each; every dz2[k][b] is still one fma chain over the logits, ascending)
        if (tid >= UNT / 2) {
          head_wgrad(sm.Lg, sm.H2, p.sp.L, part + p.lo.w_act, first, tid - UNT / 2);
          row_sums(sm.Lg, p.sp.L, part + p.lo.b_act, first, tid, UNT - MAXL);
        } else {
          const int kb = (tid >> 7) * (HID / 2);
          float acc[HID / 2];
#pragma unroll
          for (int k = 0; k < HID / 2; ++k) acc[k] = 0.f;
          for (int l = 0; l < p.sp.L; ++l) {
            const float d = sm.Lg[l * LDA + sb];
#pragma unroll
            for (int k0 = 0; k0 < HID / 2; k0 += 4) {
              const float4 w = *reinterpret_cast<const float4*>(sm.pol.w_act + l * LDW + kb + k0);
              acc[k0 + 0] = fmaf(w.x, d, acc[k0 + 0]);
              acc[k0 + 1] = fmaf(w.y, d, acc[k0 + 1]);
              acc[k0 + 2] = fmaf(w.z, d, acc[k0 + 2]);
              acc[k0 + 3] = fmaf(w.w, d, acc[k0 + 3]);
            }
          }
#pragma unroll
          for (int k = 0; k < HID / 2; ++k) {
            const float h = sm.H2[(kb + k) * LDA + sb];
            sm.D1[(kb + k) * LDA + sb] = acc[k] * (1.0f - h * h);
          }
        }
        PTH_PROF(6);  // head wgrad | dz2
        tower_backward<BOX, ADAP>(p, sm, Xs, sm.H1, sm.pol.w_pi1, part + p.lo.w_pi0, part + p.lo.b_pi0,
                                  part + p.lo.w_pi1, part + p.lo.b_pi1, nb, first, tid, prof_last, c, 18, Cx);
        PTH_PROF(7);  // pi tower backward (wgrad64, backprop64, first-layer gradient)

        // ================= value tower (a context tile has none: a CTA whose first tile is one
        // stores zeros for the value tower's and the value head's sums)
        if constexpr (ADAP) {
          if (ctile) {
            if (first) {
              for (int i = p.lo.w_vf0 + tid; i < p.lo.w_act; i += UNT) part[i] = 0.f;
              for (int i = p.lo.w_val + tid; i < P; i += UNT) part[i] = 0.f;
            }
            first = false;
            continue;
          }
        }
        __syncthreads();
        if constexpr (BOX) {
          first_layer_box<true, UNT, BT, !ADAP>(p.sp.F, Xs, p.params + p.lo.w_vf0, sm.pol.b_vf0, sm.H1, tid);
          __syncthreads();
          if constexpr (ADAP) {
            context_columns_tanh<true>(p.C, Cx, p.params + p.lo.w_vf0 + p.sp.F * HID, sm.H1, tid >> 5, UNT / 32,
                                       tid & 31);
            __syncthreads();
          }
        }
        PTH_PROF(8);  // vf first layer (one-hot: done with the policy tower's)
        dense64<true, UNT / 2, BT / 2, LDA>(V1 + (tid >> 8) * (BT / 2), sm.pol.w_vf1, sm.pol.b_vf1,
                                             sm.H2 + (tid >> 8) * (BT / 2), tid & (UNT / 2 - 1));
        __syncthreads();
        PTH_PROF(9);  // vf hidden layer
        if (lane) {
          const float v = value_head(sm.H2, sm.pol, tid);
          const float dret = ret - v;
          s_v = valid ? dret * dret : 0.f;
          const float dv = valid ? ((p.vf_coef * 2.0f) * (v - ret)) * invB : 0.f;
          sm.Lg[tid] = dv;
#pragma unroll
          for (int k = 0; k < HID; ++k) {
            const float h = sm.H2[k * LDA + tid];
            sm.D1[k * LDA + tid] = (sm.pol.w_val[k] * dv) * (1.0f - h * h);
          }
        }
        __syncthreads();  // dv vector + D1 complete
        if (tid >= BT && tid < BT + HID) {
          const int k = tid - BT;
          float acc = 0.f;
          for (int b0 = 0; b0 < BT; b0 += 4) {
            const float4 d = *reinterpret_cast<const float4*>(sm.Lg + b0);
            const float4 h = *reinterpret_cast<const float4*>(sm.H2 + k * LDA + b0);
            acc = fmaf(d.x, h.x, acc);
            acc = fmaf(d.y, h.y, acc);
            acc = fmaf(d.z, h.z, acc);
            acc = fmaf(d.w, h.w, acc);
          }
          acc_store(part + p.lo.w_val + k, acc, first);
        } else if (tid == BT + HID) {
          float s = 0.f;
          for (int b = 0; b < BT; ++b) s = s + sm.Lg[b];
          acc_store(part + p.lo.b_val, s, first);
        }
        PTH_PROF(10);  // value head + its gradients
        tower_backward<BOX, ADAP>(p, sm, Xs, V1, sm.pol.w_vf1, part + p.lo.w_vf0, part + p.lo.b_vf0,
                                  part + p.lo.w_vf1, part + p.lo.b_vf1, nb, first, tid, prof_last, c, 20, Cx);
        PTH_PROF(11);  // vf tower backward

        // ---- tile statistics: the five 128-lane trees of the contract in one pass (xor tree inside
        // each of the first four warps, then the four warp sums left to right)
        {
          float sv[5] = {s_pl, s_v, s_e, s_kl, s_cf};
#pragma unroll
          for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sv[i] = sv[i] + __shfl_xor_sync(0xffffffffu, sv[i], d);
          __syncthreads();  // sm.bc may still be read (norm partials of the previous minibatch)
          if ((tid & 31) == 0 && tid < BT)
#pragma unroll
            for (int i = 0; i < 5; ++i) sm.bc[i * 4 + (tid >> 5)] = sv[i];
          __syncthreads();
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float ts = ((sm.bc[i * 4] + sm.bc[i * 4 + 1]) + sm.bc[i * 4 + 2]) + sm.bc[i * 4 + 3];
            cta_stat[i] = first ? ts : cta_stat[i] + ts;
          }
        }
        first = false;
      }
      if (tid < 5) p.stat_part[c * 8 + tid] = cta_stat[tid];
      if constexpr (ADAP || MOD) {
        if (tid == 0) p.stat_part[c * 8 + 5] = cta_ctx;
      }
      PTH_PROF(12);  // tile statistics
      grid.sync();  // ---------------------------------------------- (1) partials written
      PTH_PROF(13);  // barrier 1 (includes waiting for the slowest CTA)

      // ---- ordered reduction of this CTA's parameter slice + squared-norm partial.
      // Ordered reduction of this CTA's parameter slice.  A partial sums per parameter (A = CTAs that
      // had a tile) are fetched RH at a time per thread, all of them in flight at once; with
      // A <= RH one thread adds a parameter's chain alone (512 parameters per pass), otherwise 2 or
      // 4 thread groups split the chain, the lower group first, handing the running sum up through
      // shared memory — the order of the additions is the contract's CTA order.  The finished sums
      // are parked in shared memory, where the squared-norm lanes (t < 128 owns parameters t,
      // t + 128, ... of the slice in ascending order) pick them up.
      float q = 0.f;
      const uint32_t epoch = p.flag_epoch + (uint32_t)id + 1u;  // tag of this minibatch's exchange
      const int par = (int)(epoch & 1u);  // alternates across launches too (flag_epoch is monotonic)
      {
        constexpr int RH = 64;  // max co-resident CTAs is 160 < 4 * RH
        const int lg = A <= RH ? 0 : (A <= 2 * RH ? 1 : 2);
        const int ngp = 1 << lg, PL = UNT >> lg;  // thread groups along the chain, parameters per pass
        const int grp = tid >> (9 - lg), li = tid & (PL - 1);
        static_assert(UNT == 512, "group index uses log2(UNT) = 9");
        const int cc0 = grp * RH;
        float* gs = sm.H1;  // PL sums (H1 is free between tiles)
        for (int i0c = 0; i0c < S; i0c += PL) {
          const int i = i0c + li;
          const int pi = c * S + i;
          bool live = i < S && pi < P;
          if constexpr (MOD) {  // value modules of the partners that are not being trained: no gradient
            const ModBlock mbk = mod_block(p.sp.L);
            live = live && mod_group(pi, p.lo.total, mbk.total, mbk.w_vf0, mbk.w_act, mbk.w_val, p.p0) != 2;
          }
          // (one base pointer + a 32-bit row offset per load: the 64-bit index arithmetic of
          // `part[(cc0 + u) * PS + pi]` was ~10 instructions per load, profiles/update_r02_hot_lines.txt)
          float t[RH];
          const float* src = p.part + (size_t)cc0 * PS + pi;
          const int nlive = live ? (A - cc0 < RH ? A - cc0 : RH) : 0;  // partial sums this thread group adds
#pragma unroll
          for (int u = 0; u < RH; ++u) t[u] = u < nlive ? __ldcg(src + u * PS) : 0.f;
          float g = 0.f;
#pragma unroll 1
          for (int ph = 0; ph < ngp; ++ph) {
            if (grp == ph) {
              if (live && cc0 < A) {
                g = ph == 0 ? t[0] : gs[li] + t[0];
#pragma unroll
                for (int u = 1; u < RH; ++u) g = u < nlive ? g + t[u] : g;
                gs[li] = g;
              } else if (ph == 0) {
                gs[li] = 0.f;  // beyond the slice, or no tile at all this minibatch
              }
            }
            __syncthreads();
          }
          if (grp == 0 && live) {
            g = gs[li];
            if (W == 1) {
              p.grad[pi] = g;
            } else {
              // this rank's ordered sum goes to every rank's exchange slot [par][rank] over NVLink
              for (int k = 0; k < W; ++k) {
                const int dst = (p.rank + k) % W;
                ll_store(p.xbuf[dst] + ((size_t)par * W + p.rank) * p.XS + pi, g, epoch);
              }
            }
          }
          if (W == 1 && tid < BT)
            for (int k = 0; k < PL / BT; ++k) {
              const float gg = gs[tid + k * BT];
              q = fmaf(gg, gg, q);  // 0 beyond the slice: q unchanged
            }
          __syncthreads();  // gs is rewritten by the next pass
        }
      }
      // loss statistics of the minibatch: one warp of the last CTA fetches every CTA's 5 sums in
      // one L2 round trip, parks them in shared memory, and 5 lanes add them in CTA order —
      // off the critical path of the parameter update (read back after barrier 2)
      if (c == G - 1 && (tid >> 5) == 4) {
        const int ln = tid & 31;
        constexpr int NST = (ADAP || MOD) ? 6 : 5;  // ADAP: + the context-loss sum; MOD: + the marginal regulariser's
        float* sc = sm.Lg;  // [NST][160] scratch (free between tiles)
        float t[NST][5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const int cc = r * 32 + ln;
#pragma unroll
          for (int i = 0; i < NST; ++i) t[i][r] = cc < A ? __ldcg(p.stat_part + cc * 8 + i) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
          for (int i = 0; i < NST; ++i) sc[i * 160 + r * 32 + ln] = t[i][r];
        __syncwarp();
        if (ln < NST) {
          float s_ = 0.f;
          if (A > 0) {
            s_ = sc[ln * 160];
            for (int cc = 1; cc < A; ++cc) s_ = s_ + sc[ln * 160 + cc];
          }
          if (W == 1) {
            p.stat_part[G * 8 + ln] = s_;
          } else {  // this rank's sums go to every rank's exchange slot, behind the gradient slice
            for (int k = 0; k < W; ++k)
              ll_store(p.xbuf[(p.rank + k) % W] + ((size_t)par * W + p.rank) * p.XS + P + ln, s_, epoch);
          }
        }
      }
      if (W > 1) {
        // rank-order sum of the slice: one parameter per thread, all ranks' values polled at once;
        // the squared-norm lanes (t < 128 owns parameters t, t + 128, ... in ascending order) pick the
        // sums up from shared memory
        const uint2* xl = p.xbuf[p.rank] + (size_t)par * W * p.XS;
        float* gs = sm.H1;
        for (int i0c = 0; i0c < S; i0c += UNT) {
          const int i = i0c + tid;
          const int pi = c * S + i;
          const bool live = i < S && pi < P;
          uint2 raw[8];
#pragma unroll
          for (int r = 0; r < 8; ++r)
            raw[r] = (live && r < W) ? ll_peek(xl + (size_t)r * p.XS + pi) : make_uint2(0u, epoch);
          float t[8];
#pragma unroll
          for (int r = 0; r < 8; ++r)
            t[r] = (live && r < W) ? ll_wait(xl + (size_t)r * p.XS + pi, epoch, raw[r]) : 0.f;
          float g = t[0];
#pragma unroll
          for (int r = 1; r < 8; ++r) g = r < W ? g + t[r] : g;  // rank order
          if (live) p.grad[pi] = g;
          gs[tid] = live ? g : 0.f;
          __syncthreads();
          if (tid < BT) {
#pragma unroll
            for (int k = 0; k < UNT / BT; ++k) {
              const float gg = gs[tid + k * BT];
              q = fmaf(gg, gg, q);
            }
          }
          __syncthreads();
        }
      }
      const float sq = block_tree(q, sm.red, tid);
      if (tid == 0) p.norm_part[c] = sq;
      PTH_PROF(14);  // ordered reduction of the slice
      grid.sync();  // ---------------------------------------------- (2) gradient + norm partials
      PTH_PROF(15);  // barrier 2

      __syncthreads();
      for (int cc = tid; cc < G; cc += UNT) sm.bc[cc] = __ldcg(p.norm_part + cc);
      // this thread's first parameter of the slice: fetched while the norm is being summed
      const int pi0 = c * S + tid;
      const bool own0 = tid < S && pi0 < P;
      float g0 = 0.f, m0 = 0.f, v0 = 0.f, w0 = 0.f;
      if (own0) {
        g0 = p.grad[pi0];
        m0 = p.adam_m[pi0];
        v0 = p.adam_v[pi0];
        w0 = p.params[pi0];
      }
      __syncthreads();
      float total_sq = sm.bc[0];
      for (int cc = 1; cc < G; ++cc) total_sq = total_sq + sm.bc[cc];
      const float gnorm = sqrtf(total_sq);
      float coef = p.max_norm / (gnorm + 1e-6f);
      if (coef > 1.0f) coef = 1.0f;
      b1pow *= (double)p.b1;
      b2pow *= (double)p.b2;
      const float step_size = (float)((double)p.lr / (1.0 - b1pow));
      const float bc2_sqrt = (float)sqrt(1.0 - b2pow);
      [[maybe_unused]] float step_size_v = 0.f, bc2_sqrt_v = 1.f;
      if constexpr (MOD) {  // the trained partner's value modules count their own optimiser steps
        b1pow_v *= (double)p.b1;
        b2pow_v *= (double)p.b2;
        step_size_v = (float)((double)p.lr / (1.0 - b1pow_v));
        bc2_sqrt_v = (float)sqrt(1.0 - b2pow_v);
      }
      for (int i = tid; i < S; i += UNT) {
        const int pi = c * S + i;
        if (pi < P) {
          float ss = step_size, bq = bc2_sqrt;
          if constexpr (MOD) {
            const ModBlock mbk = mod_block(p.sp.L);
            const int gi = mod_group(pi, p.lo.total, mbk.total, mbk.w_vf0, mbk.w_act, mbk.w_val, p.p0);
            if (gi == 2) continue;
            if (gi == 1) {
              ss = step_size_v;
              bq = bc2_sqrt_v;
            }
          }
          const bool pre = i == tid && own0;
          float g = (pre ? g0 : p.grad[pi]) * coef;
          if (p.l2 != 0.f) g = fmaf(p.l2, pre ? w0 : p.params[pi], g);  // d/dtheta of l2 * sum(theta^2) / 2
          const float mm = fmaf(omb1, g, p.b1 * (pre ? m0 : p.adam_m[pi]));
          const float vv = fmaf(omb2 * g, g, p.b2 * (pre ? v0 : p.adam_v[pi]));
          const float denom = sqrtf(vv) / bq + p.eps;
          p.adam_m[pi] = mm;
          p.adam_v[pi] = vv;
          p.params[pi] = fmaf(-ss, mm / denom, pre ? w0 : p.params[pi]);
        }
      }
      if (c == 0 && tid < 32 && p.stats) {
        // warp 0 of CTA 0 writes the minibatch's logged scalars.  Multi-GPU: lane (r, i) polls rank r's
        // i-th sum — all W x 5 pairs in flight at once (one thread polling them one after the other cost
        // 40 dependent L2 round trips = 10 us per minibatch at 8 GPUs, with every CTA of the grid waiting
        // for it at barrier 3: profiles/scaling_r02.md) — then lanes 0..4 add them in rank order.
        float mine = 0.f;
        if (W == 1) {
          if (tid < ((ADAP || MOD) ? 6 : 5)) mine = __ldcg(p.stat_part + G * 8 + tid);  // summed in CTA order during the reduce phase
        } else {
          const uint2* xl = p.xbuf[p.rank] + (size_t)par * W * p.XS + P;
          float* sc = sm.Lg;  // [8][5] scratch (free between tiles)
          for (int idx = tid; idx < W * 5; idx += 32) {
            const uint2* src = xl + (size_t)(idx / 5) * p.XS + (idx % 5);
            sc[idx] = ll_wait(src, epoch, ll_peek(src));
          }
          __syncwarp();
          if (tid < 5) {
            mine = sc[tid];
            for (int r = 1; r < W; ++r) mine = mine + sc[r * 5 + tid];
          }
        }
        float st[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) st[i] = __shfl_sync(0xffffffffu, mine, i);
        if (tid == 0) {
          float* o = p.stats + 8 * id;
          o[0] = -(st[0] / Bf);
          o[1] = st[1] / Bf;
          o[2] = -(st[2] / Bf);
          o[3] = st[3] / Bf;
          o[4] = st[4] / Bf;
          o[5] = (o[0] + p.ent_coef * o[2]) + p.vf_coef * o[1];
          o[6] = gnorm;
          o[7] = Bf;
        }
        if constexpr (ADAP) {
          const float csum = __shfl_sync(0xffffffffu, mine, 5);
          if (tid == 0) {
            float cl = 0.f;
            if (n_ctx_tiles > 0) {
              cl = (csum / (float)(p.K * (p.K - 1) / 2)) / (float)S_eff;
              p.stats[8 * id + 5] = p.stats[8 * id + 5] + p.ctx_coeff * cl;
            }
            if (p.ctx_loss) p.ctx_loss[id] = cl;
          }
        }
        if constexpr (MOD) {
          const float csum = __shfl_sync(0xffffffffu, mine, 5);
          if (tid == 0) {
            const float mg = csum / Bf;  // mean_b sum_l |main marginal - composed marginal|
            p.stats[8 * id + 5] = p.stats[8 * id + 5] + p.marg_coef * mg;
            if (p.ctx_loss) p.ctx_loss[id] = mg;
          }
        }
      }
      PTH_PROF(16);  // clip + Adam
      grid.sync();  // ---------------------------------------------- (3) parameters updated
      PTH_PROF(17);  // barrier 3
      __threadfence();  // invalidate L1: the next minibatch gathers fresh first-layer rows through L1
    }
  }
}

// ----------------------------------------------------------------- helpers
__device__ __forceinline__ uint32_t feistel(uint32_t x, int half_bits, const pth_u4& key) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t l = x >> half_bits, r = x & mask;
  const uint32_t k[4] = {key.x, key.y, key.z, key.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t f = (r ^ k[i]) * 0x9E3779B1u;
    f ^= f >> 15;
    f *= 0x85EBCA77u;
    f ^= f >> 13;
    const uint32_t nl = r, nr = (l ^ f) & mask;
    l = nl;
    r = nr;
  }
  return (l << half_bits) | r;
}

__global__ void perm_feistel_kernel(int32_t* perm, int64_t M, int n_epochs, uint64_t seed,
                                    uint32_t stream, uint32_t epoch0, int half_bits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (i >= M) return;
  const pth_u4 key = pth_philox(seed, stream, 0, epoch0 + (uint32_t)e, 0);
  uint32_t x = (uint32_t)i;
  do {
    x = feistel(x, half_bits, key);
  } while ((int64_t)x >= M);
  perm[(int64_t)e * M + i] = (int32_t)x;
}

// The random draws of ADAP on Philox: per minibatch id the positions of the sampled states
// (th.randperm(B)[:S], adap/util.py:106: the first S images of a keyed Feistel permutation of [0, B))
// and K contexts from SAMPLERS[sampler] (adap/util.py:42-94: 0 "l2" unit sphere, 1 "unit_square",
// 2 "positive_square", 3 "categorical" one-hot, 4 "natural_numbers" (context_size 1)).  One CTA per id.
__global__ void adap_draw_kernel(int32_t* sidx, float* draws, int64_t n_mb, int64_t M, int64_t BS, int S, int K, int C,
                                 int sampler, uint64_t seed, uint32_t stream, uint32_t index0) {
  const int64_t id = blockIdx.x;
  const uint32_t tick = index0 + (uint32_t)id;
  if (sidx != nullptr) {
    const int64_t m = id % n_mb;
    const int64_t B = (m * BS + BS <= M) ? BS : (M - m * BS);
    int bits = 2;
    while (((int64_t)1 << bits) < B) ++bits;
    if (bits & 1) ++bits;
    const pth_u4 key = pth_philox(seed, stream, 0, tick, 0);
    for (int s_ = threadIdx.x; s_ < S; s_ += blockDim.x) {
      int32_t out = -1;
      if (s_ < B) {
        uint32_t x = (uint32_t)s_;
        do {
          x = feistel(x, bits / 2, key);
        } while ((int64_t)x >= B);
        out = (int32_t)x;
      }
      sidx[id * S + s_] = out;
    }
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float u[MAX_CTX];
#pragma unroll
    for (int q = 0; q < MAX_CTX / 4; ++q) {
      const pth_u4 r = pth_philox(seed, stream, 1 + (uint64_t)k, tick, (uint32_t)q);
      u[4 * q + 0] = pth_u01(r.x);
      u[4 * q + 1] = pth_u01(r.y);
      u[4 * q + 2] = pth_u01(r.z);
      u[4 * q + 3] = pth_u01(r.w);
    }
    float* o = draws + ((size_t)id * K + k) * C;
    if (sampler <= 1) {
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < MAX_CTX; ++c) {
        u[c] = u[c] * 2.0f - 1.0f;
        if (c < C) ss = fmaf(u[c], u[c], ss);
      }
      const float nrm = sampler == 0 ? sqrtf(ss) : 1.0f;
#pragma unroll
      for (int c = 0; c < MAX_CTX; ++c)
        if (c < C) o[c] = sampler == 0 ? u[c] / nrm : u[c];
    } else if (sampler == 2) {
#pragma unroll
      for (int c = 0; c < MAX_CTX; ++c)
        if (c < C) o[c] = u[c];
    } else {
      int idx = (int)(u[0] * (float)C);
      idx = idx < C - 1 ? idx : C - 1;
      for (int c = 0; c < C; ++c) o[c] = sampler == 3 ? (c == idx ? 1.0f : 0.0f) : (c == 0 ? (float)idx : 0.0f);
    }
  }
}

// exclusive prefix of min(count, T) over envs, single CTA (N is at most a few 1e5)
__global__ void __launch_bounds__(1024) index_scan_kernel(const int32_t* count, int64_t T, int64_t N,
                                                          int32_t* offsets, int32_t* total) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < N; base += 1024) {
    const int64_t n = base + tid;
    int32_t v = 0;
    if (n < N) {
      v = count ? count[n] : (int32_t)T;
      if (v > T) v = (int32_t)T;
      if (v < 0) v = 0;
    }
    int32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int32_t w = warp_tot[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int32_t before = carry_s + (wid > 0 ? warp_tot[wid - 1] : 0) + (x - v);
    if (n < N) offsets[n] = before;
    __syncthreads();
    if (tid == 1023) carry_s = before + v;
    __syncthreads();
  }
  if (tid == 0) *total = carry_s;
}

__global__ void index_fill_kernel(const int32_t* count, const int32_t* offsets, int64_t T, int64_t N,
                                  int32_t* index) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int32_t c = count ? count[n] : (int32_t)T;
  if (c > T) c = (int32_t)T;
  const int32_t o = offsets[n];
  for (int32_t t = 0; t < c; ++t) index[o + t] = (int32_t)(t * N + n);
}

struct WsLayout {
  size_t part, grad, norm_part, stat_part, advstat, total;
};

WsLayout ws_layout(int G, int P, int64_t n_stat) {
  WsLayout w;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 255) / 256 * 256;
    return r;
  };
  w.part = take(sizeof(float) * (size_t)G * ((P + 3) & ~3));
  w.grad = take(sizeof(float) * P);
  w.norm_part = take(sizeof(float) * G);
  w.stat_part = take(sizeof(float) * (G + 1) * 8);  // + one row: the minibatch sums
  w.advstat = take(sizeof(float) * 2 * n_stat);
  w.total = o;
  return w;
}

constexpr size_t SMEM_ONEHOT = sizeof(UpdSmemT<false>) + sizeof(float) * HID * LDA;
constexpr size_t SMEM_BOX = sizeof(UpdSmemT<false>) + sizeof(float) * HID * LDA;
constexpr size_t SMEM_WIDE = sizeof(UpdSmemT<true>) + sizeof(float) * HID * LDA;
// ADAP: + the [MAX_CTX][LDA] context tile behind Xs / V1
constexpr size_t SMEM_ADAP = sizeof(UpdSmemT<false, true>) + sizeof(float) * (HID + MAX_CTX) * LDA;
static_assert(SMEM_ONEHOT <= 227 * 1024 && SMEM_WIDE <= 227 * 1024 && SMEM_ADAP <= 227 * 1024,
              "one CTA per SM: 227 KB of shared memory");

// kernel variant: 0 one-hot rows of 32 bytes, 1 Box rows, 2 one-hot rows of 96 bytes; 3 / 4: the
// AdapPolicy variants of 0 / 1 (context inputs + context-loss tiles)
const void* update_fn(int kind) {
  switch (kind) {
    case 1: return (const void*)ppo_update_kernel<true, false>;
    case 2: return (const void*)ppo_update_kernel<false, true>;
    case 3: return (const void*)ppo_update_kernel<false, false, true>;
    case 4: return (const void*)ppo_update_kernel<true, false, true>;
    case 5: return (const void*)ppo_update_kernel<false, false, false, true>;  // ModularAlgorithm, one-hot
    case 6: return (const void*)ppo_update_kernel<true, false, false, true>;   // ModularAlgorithm, Box
    case 7: return (const void*)ppo_update_kernel<false, false, true, false, true>;  // AdapPolicyMult, one-hot
    case 8: return (const void*)ppo_update_kernel<true, false, true, false, true>;   // AdapPolicyMult, Box
    default: return (const void*)ppo_update_kernel<false, false>;
  }
}
size_t update_smem(int kind) {
  return kind == 1 ? SMEM_BOX : (kind == 2 ? SMEM_WIDE : (kind >= 3 ? SMEM_ADAP : SMEM_ONEHOT));
}

int max_coop_ctas(const pth_ctx* ctx, int kind = 0) {
  static int cached[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
  if (cached[kind] < 0) {
    const void* fn = update_fn(kind);
    const size_t smem = update_smem(kind);
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, UNT, smem) != cudaSuccess) per_sm = 0;
    cached[kind] = per_sm * ctx->sm_count;
  }
  return cached[kind];
}

int space_is_box(const pth_space* sp) {  // the kernel variant of a space
  if (sp && sp->obs_kind == PTH_OBS_BOX) return 1;
  return sp && sp->obs_len > 32 ? 2 : 0;
}

int auto_grid(const pth_ctx* ctx, int64_t M, int64_t BS, int world = 1, int box = 0) {
  const int64_t eff = BS < M ? BS : M;
  int64_t tiles = (eff + BT - 1) / BT;
  tiles = (tiles + world - 1) / world;  // a rank computes every world-th tile
  int cap = max_coop_ctas(ctx, box);
  if (cap < 1) return 0;
  int64_t g = tiles < cap ? tiles : cap;
  // Small batches (the reference's own n_envs = 1, batch_size = 64 is ONE tile): CTAs without a
  // tile still take a slice of the ordered reduction and of Adam, which otherwise one CTA would
  // walk alone (44 k parameters: ~1 ms per minibatch instead of ~0.1 ms).
  const int64_t gmin = cap < 96 ? cap : 96;
  if (g < gmin) g = gmin;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

static long long* g_prof_buf = nullptr;
extern "C" int pth_debug_update_profile(void* d_clock_sums) {
  g_prof_buf = reinterpret_cast<long long*>(d_clock_sums);
  return PTH_OK;
}

extern "C" int64_t pth_update_xbuf_bytes(const pth_space* sp, int32_t world) {
  const int64_t P = pth_policy_param_count(sp);
  if (P < 0 || world < 1 || world > 8) return PTH_EINVAL;
  const int64_t XS = (P + 8 + 3) / 4 * 4;
  return 2 * (int64_t)world * XS * (int64_t)sizeof(uint2);  // (value, epoch tag) pairs
}

extern "C" int pth_update_grid(const pth_ctx* ctx, const pth_space* sp, int64_t M,
                               int64_t batch_size) {
  if (!ctx || M <= 0 || batch_size <= 0) return PTH_EINVAL;
  return auto_grid(ctx, M, batch_size, 1, space_is_box(sp));
}

extern "C" int64_t pth_adap_workspace_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t context_size,
                                            int64_t M, int64_t batch_size) {
  if (!ctx || !sp || M <= 0 || batch_size <= 0 || context_size < 0 || context_size > MAX_CTX) return PTH_EINVAL;
  const int64_t P = pth_adap_param_count(sp, context_size);
  if (P < 0) return PTH_EINVAL;
  const int kind = context_size > 0 ? (sp->obs_kind == PTH_OBS_BOX ? 4 : 3) : space_is_box(sp);
  const int cap = max_coop_ctas(ctx, kind);  // worst case grid (tests may pin any G <= cap)
  const int64_t n_mb = (M + batch_size - 1) / batch_size;
  return (int64_t)ws_layout(cap > 0 ? cap : 1, (int)P, 64 * n_mb).total;
}

extern "C" int64_t pth_update_workspace_bytes(const pth_ctx* ctx, const pth_space* sp, int64_t M,
                                              int64_t batch_size) {
  return pth_adap_workspace_bytes(ctx, sp, 0, M, batch_size);
}

extern "C" int64_t pth_modular_param_count(const pth_space* sp, int32_t num_partners) {
  const int64_t P = pth_policy_param_count(sp);
  if (P < 0 || num_partners < 1 || num_partners > MOD_MAX_PARTNERS) return PTH_EINVAL;
  return P + (int64_t)num_partners * mod_block(pth_space_logit_dim(sp)).total;
}

extern "C" int64_t pth_modular_workspace_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t num_partners,
                                               int64_t M, int64_t batch_size) {
  if (!ctx || !sp || M <= 0 || batch_size <= 0) return PTH_EINVAL;
  const int64_t P = pth_modular_param_count(sp, num_partners);
  if (P < 0) return PTH_EINVAL;
  const int cap = max_coop_ctas(ctx, sp->obs_kind == PTH_OBS_BOX ? 6 : 5);
  const int64_t n_mb = (M + batch_size - 1) / batch_size;
  return (int64_t)ws_layout(cap > 0 ? cap : 1, (int)P, 64 * n_mb).total;
}

extern "C" int64_t pth_adap_mult_param_count(const pth_space* sp, int32_t context_size) {
  const int64_t P = pth_policy_param_count(sp);
  if (P < 0 || context_size < 1 || context_size > MAX_CTX) return PTH_EINVAL;
  return P + 2 * ((int64_t)HID * context_size * HID + HID * context_size);
}

extern "C" int64_t pth_adap_mult_workspace_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t context_size,
                                                 int64_t M, int64_t batch_size) {
  if (!ctx || !sp || M <= 0 || batch_size <= 0) return PTH_EINVAL;
  const int64_t P = pth_adap_mult_param_count(sp, context_size);
  if (P < 0) return PTH_EINVAL;
  const int cap = max_coop_ctas(ctx, sp->obs_kind == PTH_OBS_BOX ? 8 : 7);
  const int64_t n_mb = (M + batch_size - 1) / batch_size;
  return (int64_t)ws_layout(cap > 0 ? cap : 1, (int)P, 64 * n_mb).total;
}

extern "C" int64_t pth_adap_mult_scratch_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t context_size) {
  if (!ctx || !sp || context_size < 1 || context_size > MAX_CTX) return PTH_EINVAL;
  const int cap = max_coop_ctas(ctx, sp->obs_kind == PTH_OBS_BOX ? 8 : 7);
  return (int64_t)(sizeof(float) * mult_scratch_floats(context_size)) * (cap > 0 ? cap : 1);
}

// per-CTA activation scratch of the modular tile for the largest grid a launch may use
extern "C" int64_t pth_modular_scratch_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t num_partners) {
  if (!ctx || !sp || num_partners < 1 || num_partners > MOD_MAX_PARTNERS) return PTH_EINVAL;
  const int cap = max_coop_ctas(ctx, sp->obs_kind == PTH_OBS_BOX ? 6 : 5);
  return (int64_t)(sizeof(float) * mod_scratch_floats(num_partners)) * (cap > 0 ? cap : 1);
}

extern "C" int pth_ppo_update(pth_ctx* ctx, const pth_update_args* a, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && a != nullptr, "NULL ctx/args");
  PTH_CHECK_ARG(a->space && a->d_params && a->d_adam_m && a->d_adam_v, "NULL space/params/adam");
  PTH_CHECK_ARG((a->d_obs || a->d_obs_f32) && a->d_actions && a->d_old_logp && a->d_advantages &&
                    a->d_returns && a->d_perm && a->d_workspace,
                "NULL sample array / perm / workspace");
  PTH_CHECK_ARG(a->M > 0 && a->batch_size > 0 && a->n_epochs > 0 && a->n_epochs <= 64,
                "bad M / batch_size / n_epochs");
  PTH_CHECK_ARG(a->M < ((int64_t)1 << 31), "buffer too large for int32 indices");
  UpdParams p;
  if (fill_space(a->space, &p.sp) != 0) {
    pth_set_error("pth_ppo_update: unsupported space");
    return PTH_ENOSUP;
  }
  const bool box = p.sp.obs_kind == PTH_OBS_BOX;
  if (box) {
    if (p.sp.F > HID) {
      pth_set_error("pth_ppo_update: Box observations wider than 64 are not supported");
      return PTH_ENOSUP;
    }
    PTH_CHECK_ARG(a->d_obs_f32 != nullptr && a->obs_stride == HID,
                  "Box observations: d_obs_f32 rows of 64 floats");
    PTH_CHECK_ARG(((uintptr_t)a->d_obs_f32 % 16) == 0, "obs rows must be 16-byte aligned");
  } else {
    PTH_CHECK_ARG(a->d_obs != nullptr, "NULL d_obs");
    PTH_CHECK_ARG(p.sp.F <= MAX_ROWS, "one-hot feature width above 1024 is not supported");
  }
  // AdapPolicy (pantheonrl/algos/adap/policies.py:71-131): context_size inputs behind the features
  const int C = a->context_size;
  PTH_CHECK_ARG(C >= 0 && C <= MAX_CTX, "context_size above 8 is not supported");
  PTH_CHECK_ARG(a->loss_kind != PTH_LOSS_ADAP || C > 0, "ADAP needs context_size > 0");
  PTH_CHECK_ARG(C == 0 || (a->d_context != nullptr && a->rec_stride == 0 && a->world <= 1 && (box || p.sp.obs_len <= 32)),
                "AdapPolicy: d_context rows, separate sample arrays, one GPU, at most 32 observation slots");
  p.C = C;
  p.ctx = a->d_context;
  p.ctx_coeff = 0.f;
  p.K = p.NS = 0;
  p.ctx_sidx = nullptr;
  p.ctx_draws = nullptr;
  p.ctx_loss = a->d_ctx_loss;
  if (a->loss_kind == PTH_LOSS_ADAP && a->context_loss_coeff != 0.f && a->num_context_samples >= 2) {
    PTH_CHECK_ARG(a->num_context_samples <= 16 && a->num_state_samples >= 1 && a->d_ctx_states && a->d_ctx_draws,
                  "ADAP context loss: 2..16 context samples, >= 1 state samples, d_ctx_states, d_ctx_draws");
    p.ctx_coeff = a->context_loss_coeff;
    p.K = a->num_context_samples;
    p.NS = a->num_state_samples;
    p.ctx_sidx = a->d_ctx_states;
    p.ctx_draws = a->d_ctx_draws;
  }
  p.lo = make_layout(p.sp.F + C, p.sp.L);
  p.mult_w[0] = p.mult_w[1] = p.mult_b[0] = p.mult_b[1] = 0;
  const bool mult = a->adap_mult != 0;
  if (mult) {  // AdapPolicyMult: per tower  first layer (features only) | scaling 64 -> 64 C | second layer
    PTH_CHECK_ARG(C > 0, "AdapPolicyMult needs context_size > 0");
    int o = 0;
    const int F = p.sp.F, L = p.sp.L;
    p.lo.w_pi0 = o; o += HID * F;
    p.lo.b_pi0 = o; o += HID;
    p.mult_w[0] = o; o += HID * C * HID;
    p.mult_b[0] = o; o += HID * C;
    p.lo.w_pi1 = o; o += HID * HID;
    p.lo.b_pi1 = o; o += HID;
    p.lo.w_vf0 = o; o += HID * F;
    p.lo.b_vf0 = o; o += HID;
    p.mult_w[1] = o; o += HID * C * HID;
    p.mult_b[1] = o; o += HID * C;
    p.lo.w_vf1 = o; o += HID * HID;
    p.lo.b_vf1 = o; o += HID;
    p.lo.w_act = o; o += L * HID;
    p.lo.b_act = o; o += L;
    p.lo.w_val = o; o += HID;
    p.lo.b_val = o; o += 1;
    p.lo.total = o;
  }
  p.P = p.lo.total;
  // ModularAlgorithm (pantheonrl/algos/modular): num_partners modules behind the main network
  const bool modular = a->loss_kind == PTH_LOSS_MODULAR;
  p.Pn = p.p0 = 0;
  p.marg_coef = 0.f;
  p.mod_ws = nullptr;
  p.b1pow0_v = p.b2pow0_v = 1.0;
  if (modular) {
    PTH_CHECK_ARG(C == 0 && a->rec_stride == 0 && a->world <= 1 && (box || p.sp.obs_len <= 32),
                  "ModularAlgorithm: no context inputs, separate sample arrays, one GPU, at most 32 observation slots");
    PTH_CHECK_ARG(a->num_partners >= 1 && a->num_partners <= MOD_MAX_PARTNERS && a->partner_idx >= 0 &&
                      a->partner_idx < a->num_partners,
                  "ModularAlgorithm: 1..8 partners, partner_idx among them");
    p.Pn = a->num_partners;
    p.p0 = a->partner_idx;
    p.marg_coef = a->marginal_reg_coef;
    p.P = p.lo.total + p.Pn * mod_block(p.sp.L).total;
    p.b1pow0_v = pow((double)a->adam_beta1, (double)a->partner_vf_step);
    p.b2pow0_v = pow((double)a->adam_beta2, (double)a->partner_vf_step);
    p.ctx_loss = a->d_ctx_loss;
  }
  for (int i = 0; i < MAX_SLOTS; ++i)
    p.nvec[i] = (!box && i < p.sp.obs_len) ? (uint8_t)a->space->obs_nvec[i] : 0;
  const int kind = modular ? (box ? 6 : 5)
                           : (mult ? (box ? 8 : 7) : (C > 0 ? (box ? 4 : 3) : (box ? 1 : (p.sp.obs_len > 32 ? 2 : 0))));
  const int cap = max_coop_ctas(ctx, kind);
  PTH_CHECK_ARG(cap <= 160, "more than 160 co-resident CTAs are not supported");
  if (cap < 1 || !ctx->coop_launch) {
    pth_set_error("pth_ppo_update: cooperative launch unavailable on this device");
    return PTH_ENOSUP;
  }
  int G = a->grid_ctas > 0 ? a->grid_ctas
                           : auto_grid(ctx, a->M, a->batch_size, a->world > 1 ? a->world : 1, kind);
  PTH_CHECK_ARG(G >= 1 && G <= cap, "grid_ctas exceeds the co-resident CTA capacity");
  const int64_t n_mb = (a->M + a->batch_size - 1) / a->batch_size;
  const WsLayout w = ws_layout(G, p.P, (int64_t)a->n_epochs * n_mb);
  if (modular) {
    PTH_CHECK_ARG(a->d_modular_scratch != nullptr && ((uintptr_t)a->d_modular_scratch % 16) == 0 &&
                      a->modular_scratch_bytes >= (int64_t)(sizeof(float) * mod_scratch_floats(p.Pn) * (size_t)G),
                  "ModularAlgorithm: d_modular_scratch too small (pth_modular_scratch_bytes)");
    p.mod_ws = reinterpret_cast<float*>(a->d_modular_scratch);
  }
  if (mult) {
    PTH_CHECK_ARG(a->d_modular_scratch != nullptr && ((uintptr_t)a->d_modular_scratch % 16) == 0 &&
                      a->modular_scratch_bytes >= (int64_t)(sizeof(float) * mult_scratch_floats(C) * (size_t)G),
                  "AdapPolicyMult: d_modular_scratch too small (pth_adap_mult_scratch_bytes)");
    p.mod_ws = reinterpret_cast<float*>(a->d_modular_scratch);
  }
  PTH_CHECK_ARG((int64_t)w.total <= a->workspace_bytes, "workspace too small");
  PTH_CHECK_ARG(((uintptr_t)a->d_workspace % 256) == 0 && ((uintptr_t)a->d_params % 16) == 0,
                "workspace must be 256-byte aligned, params 16-byte aligned");

  p.params = a->d_params;
  p.adam_m = a->d_adam_m;
  p.adam_v = a->d_adam_v;
  p.obs = box ? reinterpret_cast<const uint8_t*>(a->d_obs_f32) : a->d_obs;
  p.actions = a->d_actions;
  p.old_logp = reinterpret_cast<const uint8_t*>(a->d_old_logp);
  p.adv = reinterpret_cast<const uint8_t*>(a->d_advantages);
  p.ret = reinterpret_cast<const uint8_t*>(a->d_returns);
  if (a->rec_stride > 0) {
    PTH_CHECK_ARG(a->rec_stride % 16 == 0, "rec_stride must be a multiple of 16");
    p.obs_stride = p.act_stride = p.f_stride = a->rec_stride;
  } else {
    p.obs_stride = box ? (int64_t)HID * 4 : PTH_OBS_ROW_BYTES(p.sp.obs_len);
    p.act_stride = 4;
    p.f_stride = 4;
  }
  PTH_CHECK_ARG(((uintptr_t)p.obs % 16) == 0, "obs rows must be 16-byte aligned");
  p.index = a->d_index;
  p.perm = a->d_perm;
  p.M = a->M;
  p.BS = a->batch_size;
  p.n_epochs = a->n_epochs;
  p.lr = a->learning_rate;
  p.clip = a->clip_range;
  p.ent_coef = a->ent_coef;
  p.vf_coef = a->vf_coef;
  p.max_norm = a->max_grad_norm;
  p.b1 = a->adam_beta1;
  p.b2 = a->adam_beta2;
  p.eps = a->adam_eps;
  p.normalize = a->normalize_advantage;
  // ADAP = PPO's losses + context tiles; ModularAlgorithm = PPO's losses on the composed outputs + the regulariser
  p.loss_kind = (a->loss_kind == PTH_LOSS_ADAP || modular) ? PTH_LOSS_PPO : a->loss_kind;
  p.l2 = a->l2_weight;
  PTH_CHECK_ARG(a->loss_kind >= PTH_LOSS_PPO && a->loss_kind <= PTH_LOSS_MODULAR, "bad loss_kind");
  PTH_CHECK_ARG(a->loss_kind == PTH_LOSS_PPO || a->world <= 1, "behaviour cloning / ADAP run on one GPU");
  p.b1pow0 = pow((double)a->adam_beta1, (double)a->adam_step);
  p.b2pow0 = pow((double)a->adam_beta2, (double)a->adam_step);
  unsigned char* ws = reinterpret_cast<unsigned char*>(a->d_workspace);
  p.part = reinterpret_cast<float*>(ws + w.part);
  p.grad = reinterpret_cast<float*>(ws + w.grad);
  p.norm_part = reinterpret_cast<float*>(ws + w.norm_part);
  p.stat_part = reinterpret_cast<float*>(ws + w.stat_part);
  p.advstat = reinterpret_cast<float*>(ws + w.advstat);
  p.stats = a->d_stats;
  p.prof = g_prof_buf;
  p.world = a->world > 1 ? a->world : 1;
  p.rank = p.world > 1 ? a->rank : 0;
  p.flag_epoch = a->flag_epoch;
  p.XS = (p.P + 8 + 3) / 4 * 4;
  for (int r = 0; r < 8; ++r) {
    p.xbuf[r] = nullptr;
    p.advx[r] = nullptr;
  }
  if (p.world > 1) {
    PTH_CHECK_ARG(p.world <= 8 && p.rank >= 0 && p.rank < p.world, "bad world / rank");
    PTH_CHECK_ARG(a->peer_xbuf && a->peer_flags, "NULL peer pointer arrays");
    for (int r = 0; r < p.world; ++r) {
      PTH_CHECK_ARG(a->peer_xbuf[r] && a->peer_flags[r], "NULL peer buffer");
      p.xbuf[r] = reinterpret_cast<uint2*>(a->peer_xbuf[r]);
      p.advx[r] = reinterpret_cast<uint2*>(a->peer_flags[r]);
    }
    PTH_CHECK_ARG((int64_t)a->n_epochs * n_mb * 4 <= PTH_UPDATE_FLAG_WORDS,
                  "sharded update: more minibatches per launch than the advantage-statistics exchange holds");
  }

  void* kargs[] = {(void*)&p};
  PTH_CUDA(cudaLaunchCooperativeKernel(update_fn(kind), dim3(G), dim3(UNT), kargs, update_smem(kind),
                                       (cudaStream_t)stream));
  return PTH_OK;
}

extern "C" int pth_perm_feistel(pth_ctx* ctx, int32_t* d_perm, int64_t M, int32_t n_epochs,
                                uint64_t seed, uint32_t stream_id, uint32_t epoch0, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && d_perm != nullptr, "NULL ctx/perm");
  PTH_CHECK_ARG(M > 0 && M < ((int64_t)1 << 31) && n_epochs > 0 && n_epochs <= 65535, "bad M / n_epochs");
  int bits = 2;
  while (((int64_t)1 << bits) < M) ++bits;
  if (bits & 1) ++bits;
  dim3 grid((unsigned)pth_ceil_div(M, 256), (unsigned)n_epochs);
  perm_feistel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_perm, M, n_epochs, seed, stream_id,
                                                              epoch0, bits / 2);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_adap_draw(pth_ctx* ctx, int32_t* d_states, float* d_draws, int64_t n, int64_t n_minibatches,
                             int64_t M, int64_t batch_size, int32_t num_state_samples, int32_t num_context_samples,
                             int32_t context_size, int32_t sampler, uint64_t seed, uint32_t stream_id,
                             uint32_t index0, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && d_draws != nullptr && n > 0 && n < (1 << 30), "NULL ctx / draws, bad n");
  PTH_CHECK_ARG(context_size >= 1 && context_size <= MAX_CTX && num_context_samples >= 1 && sampler >= 0 && sampler <= 4,
                "bad context_size / num_context_samples / sampler");
  PTH_CHECK_ARG(sampler != 4 || context_size == 1, "natural_numbers contexts have one column");
  PTH_CHECK_ARG(d_states == nullptr || (num_state_samples >= 1 && n_minibatches >= 1 && M > 0 && batch_size > 0),
                "bad state-sample geometry");
  adap_draw_kernel<<<(unsigned)n, 64, 0, (cudaStream_t)stream>>>(d_states, d_draws, n_minibatches, M, batch_size,
                                                                  num_state_samples, num_context_samples, context_size,
                                                                  sampler, seed, stream_id, index0);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int64_t pth_index_workspace_bytes(int64_t N) { return N > 0 ? N * 4 + 256 : 256; }

extern "C" int pth_index_build(pth_ctx* ctx, const int32_t* d_count, int64_t T, int64_t N,
                               int32_t* d_index, int32_t* d_total, void* d_workspace, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && d_index && d_total && d_workspace, "NULL pointer");
  PTH_CHECK_ARG(T > 0 && N > 0 && T * N < ((int64_t)1 << 31), "bad T / N");
  int32_t* offsets = reinterpret_cast<int32_t*>(d_workspace);
  cudaStream_t st = (cudaStream_t)stream;
  index_scan_kernel<<<1, 1024, 0, st>>>(d_count, T, N, offsets, d_total);
  index_fill_kernel<<<pth_ceil_div(N, 128), 128, 0, st>>>(d_count, offsets, T, N, d_index);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}
