/*
 * pth_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the hot path of Stanford-ILIAD/PantheonRL:
 * rollout collection -> GAE -> PPO update, as driven by
 *   pantheonrl/common/agents.py:111-203      (OnPolicyAgent.get_action/update)
 *   pantheonrl/common/multiagentenv.py:149-243, 307-327, 395-409
 *   pantheonrl/envs/rpsgym/rps.py:41-48, pantheonrl/envs/liargym/liar.py:22-102
 * and the arithmetic of stable-baselines3==1.7.0 (setup.py:17; NOT vendored in
 * the reference tree, restated from its published algorithm: SURVEY.md
 * Appendix A; in-tree near copies pantheonrl/algos/adap/adap_learn.py:229-347,
 * pantheonrl/algos/modular/policies.py:84-118,214-290,364-383; GAE recurrence
 * also at overcookedgym/human_aware_rl/baselines/baselines/ppo2/runner.py:150-165).
 *
 * PARITY STATUS: the env / routing half is pinned against traces produced by
 * the reference's own Python classes (tests/golden/make_golden.py imports
 * /root/reference).  PPO.train and GAE are pinned on the reference's own
 * in-tree copies of the SB3 routines, executed verbatim
 * (tests/golden/make_golden_sb3_intree.py runs ADAP.train of
 * pantheonrl/algos/adap/adap_learn.py:229-347 and the GAE loop of
 * ppo2/runner.py:152-164; tests/test_oracle_sb3_intree.py: <= 5e-6 / 1e-5).
 * "Parity unpinned" remains for the policy construction / sampling against a
 * real SB3 + torch 1.13.1 (not installable here); that part is cross-checked
 * against an independent torch restatement (oracle/sb3_torch.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file.  Everything is scalar fp32 with a
 * fixed evaluation order (compile with -ffp-contract=off; fused multiply-add
 * only where fmaf() is written) so the CUDA path can be compared bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define H 64
#define MAX_SLOTS 96
#define MAX_HEADS 4

typedef struct {
  int32_t obs_kind; /* 0 one-hot, 1 box */
  int32_t obs_len;
  int32_t obs_nvec[MAX_SLOTS];
  int32_t n_heads;
  int32_t head_n[MAX_HEADS];
} orc_space;

/* ---------------------------------------------------------------- RNG */
/* Philox4x32-10, Salmon et al. SC'11 (Random123). */
void orc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox(uint64_t seed, uint32_t stream, uint64_t index, uint32_t tick, uint32_t slot,
                uint32_t out[4]) {
  uint32_t ctr[4] = {(uint32_t)index, tick, slot, (uint32_t)(index >> 32)};
  uint32_t key[2] = {(uint32_t)seed ^ (stream * 0x9E3779B9u), (uint32_t)(seed >> 32)};
  orc_philox_raw(ctr, key, out);
}

static float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

enum { STREAM_ENV = 1, STREAM_EGO = 2, STREAM_ALT = 3 };

/* ---------------------------------------------------------------- math */
static float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }

float orc_expf(float x) {
  if (x < -87.0f) return 0.0f;
  if (x > 88.0f) x = 88.0f;
  float n = rintf(x * 1.44269504088896341f);
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  float z = r * r;
  float y = fmaf(p, z, r);
  y = y + 1.0f;
  int ni = (int)n;
  return y * i2f((ni + 127) << 23);
}

float orc_logf(float x) {
  int32_t bits = f2i(x);
  int e = ((bits >> 23) & 0xff) - 126;
  float m = i2f((bits & 0x007fffff) | 0x3f000000);
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = (m + m) - 1.0f;
  } else {
    m = m - 1.0f;
  }
  float z = m * m;
  float p = 7.0376836292e-2f;
  p = fmaf(p, m, -1.1514610310e-1f);
  p = fmaf(p, m, 1.1676998740e-1f);
  p = fmaf(p, m, -1.2420140846e-1f);
  p = fmaf(p, m, 1.4249322787e-1f);
  p = fmaf(p, m, -1.6668057665e-1f);
  p = fmaf(p, m, 2.0000714765e-1f);
  p = fmaf(p, m, -2.4999993993e-1f);
  p = fmaf(p, m, 3.3333331174e-1f);
  float y = (p * m) * z;
  float fe = (float)e;
  y = fmaf(fe, -2.12194440e-4f, y);
  y = fmaf(z, -0.5f, y);
  float r = m + y;
  r = fmaf(fe, 0.693359375f, r);
  return r;
}

float orc_tanhf(float x) {
  float a = fabsf(x);
  if (a > 10.0f) return copysignf(1.0f, x);
  if (a >= 0.625f) {
    float s = orc_expf(a + a);
    float t = 1.0f - 2.0f / (s + 1.0f);
    return copysignf(t, x);
  }
  float z = x * x;
  float p = -5.70498872745e-3f;
  p = fmaf(p, z, 2.06390887954e-2f);
  p = fmaf(p, z, -5.37397155531e-2f);
  p = fmaf(p, z, 1.33314422036e-1f);
  p = fmaf(p, z, -3.33332819422e-1f);
  return fmaf(p * z, x, x);
}

void orc_math_vec(int which, const float* x, float* y, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    y[i] = which == 0 ? orc_expf(x[i]) : (which == 1 ? orc_logf(x[i]) : orc_tanhf(x[i]));
}

/* ---------------------------------------------------------------- GAE */
/* SB3 RolloutBuffer.compute_returns_and_advantage, float32 throughout (the
 * PantheonRL partner path, agents.py:127-130, passes a Python bool for dones
 * so no float64 promotion occurs). [T][N] layout. */
void orc_gae(const float* rew, const float* val, const float* start, const float* last_values,
             const float* dones, float* adv, float* ret, int64_t T, int64_t N, double gamma,
             double lam) {
  const float g = (float)gamma, c = (float)(gamma * lam);
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    float last = 0.f;
    for (int64_t t = T - 1; t >= 0; --t) {
      float nnt, nv;
      if (t == T - 1) {
        nnt = 1.0f - dones[n];
        nv = last_values[n];
      } else {
        nnt = 1.0f - start[(t + 1) * N + n];
        nv = val[(t + 1) * N + n];
      }
      float t1 = g * nv;
      t1 = t1 * nnt;
      float d = (rew[t * N + n] + t1) - val[t * N + n];
      float cc = c * nnt;
      last = d + cc * last;
      adv[t * N + n] = last;
      ret[t * N + n] = last + val[t * N + n];
    }
  }
}

void orc_gae_ragged(const float* rew, const float* val, const float* start, const int32_t* count,
                    const float* last_done, float* adv, float* ret, int64_t Tcap, int64_t N,
                    double gamma, double lam) {
  const float g = (float)gamma, c = (float)(gamma * lam);
  for (int64_t n = 0; n < N; ++n) {
    int64_t cnt = count[n];
    if (cnt <= 0) continue;
    if (cnt > Tcap) cnt = Tcap;
    float last = 0.f;
    for (int64_t t = cnt - 1; t >= 0; --t) {
      float nnt, nv;
      if (t == cnt - 1) {
        nnt = 1.0f - last_done[n];
        nv = val[(cnt - 1) * N + n]; /* agents.py:127-129: self.values of the last stored step */
      } else {
        nnt = 1.0f - start[(t + 1) * N + n];
        nv = val[(t + 1) * N + n];
      }
      float t1 = g * nv;
      t1 = t1 * nnt;
      float d = (rew[t * N + n] + t1) - val[t * N + n];
      float cc = c * nnt;
      last = d + cc * last;
      adv[t * N + n] = last;
      ret[t * N + n] = last + val[t * N + n];
    }
  }
}

/* ---------------------------------------------------------------- games */
/* rps.py:41-45 */
void orc_rps_step(const int32_t* ego_a, const int32_t* alt_a, float* r_ego, float* r_alt,
                  int64_t N) {
  for (int64_t n = 0; n < N; ++n) {
    int o = (ego_a[n] - alt_a[n] + 3) % 3;
    if (o == 2) o = -1;
    r_ego[n] = (float)o;
    r_alt[n] = (float)(-o);
  }
}

/* Liar's Dice state record (same 32-byte layout as pth_liar_state). */
typedef struct {
  uint8_t hands[12];
  uint8_t hist[12]; /* face | count << 3, newest first */
  uint8_t hist_len;
  uint8_t pad[7];
} orc_liar;

/* liar.py:53-56 */
static void liar_obs(const orc_liar* s, int player, uint8_t* obs /*32*/) {
  for (int i = 0; i < 6; ++i) obs[i] = s->hands[player * 6 + i];
  for (int i = 0; i < 12; ++i) {
    if (i < s->hist_len) {
      obs[6 + 2 * i] = s->hist[i] & 7;
      obs[7 + 2 * i] = s->hist[i] >> 3;
    } else {
      obs[6 + 2 * i] = 6;
      obs[7 + 2 * i] = 0;
    }
  }
  obs[30] = obs[31] = 0;
}

/* liar.py:58-83; returns done */
static int liar_step(orc_liar* s, int player, int face, int count, float* r_ego, float* r_alt) {
  int bluff = 0;
  if (s->hist_len != 0) {
    if (count <= (s->hist[0] >> 3) || face == 6) bluff = 1;
  } else if (face == 6) {
    face = 0;
    count = 0;
  }
  if (bluff) {
    int side = s->hist[0] & 7;
    int trueans = s->hands[side] + s->hands[6 + side] - 1;
    int was_bluff = (s->hist[0] >> 3) > trueans;
    int didwin = (was_bluff == (player == 0));
    *r_ego = didwin ? 1.f : -1.f;
    *r_alt = didwin ? -1.f : 1.f;
    return 1;
  }
  memmove(s->hist + 1, s->hist, 11);
  s->hist[0] = (uint8_t)(face | (count << 3));
  s->hist_len += 1;
  *r_ego = 0.f;
  *r_alt = 0.f;
  return 0;
}

/* liar.py:97-102 + multiagentenv.py:325; draw layout documented in DESIGN.md */
static int liar_reset(orc_liar* s, uint64_t seed, uint64_t env, uint32_t tick,
                      uint32_t slot_base, float probegostart) {
  uint32_t d[16];
  for (uint32_t k = 0; k < 4; ++k) orc_philox(seed, STREAM_ENV, env, tick, slot_base + k, d + 4 * k);
  int ego_first = u01(d[0]) < probegostart;
  memset(s, 0, sizeof(*s));
  for (int i = 0; i < 12; ++i) {
    uint32_t side = (uint32_t)(((uint64_t)d[1 + i] * 6u) >> 32);
    s->hands[(i / 6) * 6 + side] += 1;
  }
  return ego_first;
}

void orc_liar_reset(orc_liar* state, uint8_t* ego_first, uint8_t* obs, int64_t N, uint64_t seed,
                    uint32_t tick, int64_t env0, float probegostart) {
  for (int64_t n = 0; n < N; ++n) {
    int ef = liar_reset(&state[n], seed, (uint64_t)(env0 + n), tick, 0, probegostart);
    if (ego_first) ego_first[n] = (uint8_t)ef;
    if (obs) liar_obs(&state[n], ef ? 0 : 1, obs + 32 * n);
  }
}

void orc_liar_step(orc_liar* state, const uint8_t* is_ego, const uint8_t* action, uint8_t* obs,
                   float* r_ego, float* r_alt, uint8_t* done, int64_t N) {
  for (int64_t n = 0; n < N; ++n) {
    int player = is_ego[n] ? 0 : 1;
    done[n] = (uint8_t)liar_step(&state[n], player, action[2 * n], action[2 * n + 1], &r_ego[n],
                                 &r_alt[n]);
    liar_obs(&state[n], 1 - player, obs + 32 * n);
  }
}

/* ---------------------------------------------------------------- policy */
static int feat_dim(const orc_space* sp) {
  if (sp->obs_kind == 1) return sp->obs_len;
  int f = 0;
  for (int s = 0; s < sp->obs_len; ++s) f += sp->obs_nvec[s];
  return f;
}
static int logit_dim(const orc_space* sp) {
  int l = 0;
  for (int h = 0; h < sp->n_heads; ++h) l += sp->head_n[h];
  return l;
}
int64_t orc_param_count(const orc_space* sp) {
  int64_t F = feat_dim(sp), L = logit_dim(sp);
  return 2 * (H * F + H + H * H + H) + L * H + L + H + 1;
}

typedef struct {
  const float *w_pi0, *b_pi0, *w_pi1, *b_pi1, *w_vf0, *b_vf0, *w_vf1, *b_vf1, *w_act, *b_act,
      *w_val, *b_val;
  int F, L;
  int C; /* ADAP (pantheonrl/algos/adap/policies.py:71-84): context inputs appended to the features;
            the first-layer matrices then have F + C rows, the context rows last */
} orc_params;

static void split_params_ctx(const orc_space* sp, const float* p, int C, orc_params* q) {
  int F = feat_dim(sp), L = logit_dim(sp);
  q->F = F; q->L = L; q->C = C;
  q->w_pi0 = p; p += H * (F + C);
  q->b_pi0 = p; p += H;
  q->w_pi1 = p; p += H * H;
  q->b_pi1 = p; p += H;
  q->w_vf0 = p; p += H * (F + C);
  q->b_vf0 = p; p += H;
  q->w_vf1 = p; p += H * H;
  q->b_vf1 = p; p += H;
  q->w_act = p; p += L * H;
  q->b_act = p; p += L;
  q->w_val = p; p += H;
  q->b_val = p;
}
static void split_params(const orc_space* sp, const float* p, orc_params* q) { split_params_ctx(sp, p, 0, q); }

/* features: SB3 preprocess_obs (one-hot concat) — Appendix A2.  A linear layer
 * is evaluated as acc = bias; for k ascending: acc = fma(x_k, w[j][k], acc).
 * One-hot first layers add the selected rows slot by slot in DESCENDING slot order
 * (x in {0,1}: an exact product, so only the order of the additions is a choice).  Descending,
 * because the games pad their observations at the END (Liar's Dice: empty history pairs): the
 * chain then begins with rows that are the same for most samples of a tile, and the update
 * kernel evaluates that common beginning once per tile instead of once per sample
 * (same additions, same bits).
 * Storage: the two first-layer matrices are kept input-major [F][64] (the
 * transpose of torch's nn.Linear.weight); every other tensor is [out][in]. */
static void first_layer(const orc_space* sp, const void* obs_row, const float* w, const float* b,
                        int F, float* out /*H, pre-activation*/) {
  if (sp->obs_kind == 0) {
    const uint8_t* o = (const uint8_t*)obs_row;
    for (int j = 0; j < H; ++j) {
      float acc = b[j];
      int off = F;
      for (int s = sp->obs_len - 1; s >= 0; --s) {
        off -= sp->obs_nvec[s];
        acc = acc + w[(off + o[s]) * H + j];
      }
      out[j] = acc;
    }
  } else {
    const float* x = (const float*)obs_row;
    for (int j = 0; j < H; ++j) {
      float acc = b[j];
      for (int k = 0; k < F; ++k) acc = fmaf(x[k], w[k * H + j], acc);
      out[j] = acc;
    }
  }
}

static void dense(const float* x, int K, const float* w, const float* b, int J, float* out) {
  for (int j = 0; j < J; ++j) {
    float acc = b[j];
    for (int k = 0; k < K; ++k) acc = fmaf(x[k], w[j * K + k], acc);
    out[j] = acc;
  }
}

typedef struct {
  float h1p[H], h2p[H], h1v[H], h2v[H], logits[32 * MAX_HEADS], value;
} orc_acts;

/* ctx: the q->C context inputs of an AdapPolicy (adap/policies.py:86-106: features = cat(features,
 * context)); they continue the first layer's chain after the features, c ascending. */
static void context_columns(const orc_params* q, const float* w, const float* ctx, float* z) {
  for (int j = 0; j < H; ++j)
    for (int c = 0; c < q->C; ++c) z[j] = fmaf(ctx[c], w[(q->F + c) * H + j], z[j]);
}

static void forward_one_ctx(const orc_space* sp, const orc_params* q, const void* obs_row, const float* ctx,
                            orc_acts* a) {
  float z[H];
  first_layer(sp, obs_row, q->w_pi0, q->b_pi0, q->F, z);
  if (q->C) context_columns(q, q->w_pi0, ctx, z);
  for (int j = 0; j < H; ++j) a->h1p[j] = orc_tanhf(z[j]);
  dense(a->h1p, H, q->w_pi1, q->b_pi1, H, z);
  for (int j = 0; j < H; ++j) a->h2p[j] = orc_tanhf(z[j]);
  dense(a->h2p, H, q->w_act, q->b_act, q->L, a->logits);
  first_layer(sp, obs_row, q->w_vf0, q->b_vf0, q->F, z);
  if (q->C) context_columns(q, q->w_vf0, ctx, z);
  for (int j = 0; j < H; ++j) a->h1v[j] = orc_tanhf(z[j]);
  dense(a->h1v, H, q->w_vf1, q->b_vf1, H, z);
  for (int j = 0; j < H; ++j) a->h2v[j] = orc_tanhf(z[j]);
  dense(a->h2v, H, q->w_val, q->b_val, 1, &a->value);
}
static void forward_one(const orc_space* sp, const orc_params* q, const void* obs_row, orc_acts* a) {
  forward_one_ctx(sp, q, obs_row, NULL, a);
}

/* One categorical head.  Inverse-CDF sampling on a single uniform (our RNG
 * contract; torch.multinomial's exponential race is not reproducible across
 * torch versions).  p_i = exp(z_i - max); S = sum ascending; pick the first i
 * with cumsum_i > u * S (last index as the fallback).
 * log_prob = (z_a - max) - log S; entropy = -sum (p_i / S) * ((z_i - max) - log S). */
static void head_eval(const float* z, int n, int sample, float u, int* action, float* logp,
                      float* entropy) {
  float m = z[0];
  for (int i = 1; i < n; ++i) m = z[i] > m ? z[i] : m;
  float p[32];
  float S = 0.f;
  for (int i = 0; i < n; ++i) {
    p[i] = orc_expf(z[i] - m);
    S = S + p[i];
  }
  float logS = orc_logf(S);
  if (sample) {
    float thr = u * S;
    float cum = 0.f;
    int a = n - 1;
    for (int i = 0; i < n; ++i) {
      cum = cum + p[i];
      if (cum > thr) { a = i; break; }
    }
    *action = a;
  }
  *logp = (z[*action] - m) - logS;
  float ent = 0.f;
  for (int i = 0; i < n; ++i) {
    float lp = (z[i] - m) - logS;
    float pi = p[i] / S;
    ent = fmaf(-pi, lp, ent);
  }
  *entropy = ent;
}

static void dist_eval(const orc_space* sp, const float* logits, int sample, const uint32_t* rnd,
                      uint8_t* action /*4*/, float* logp, float* entropy) {
  float lp = 0.f, en = 0.f;
  int off = 0;
  for (int h = 0; h < sp->n_heads; ++h) {
    int a = action[h];
    float l, e;
    head_eval(logits + off, sp->head_n[h], sample, sample ? u01(rnd[h]) : 0.f, &a, &l, &e);
    action[h] = (uint8_t)a;
    lp = lp + l;
    en = en + e;
    off += sp->head_n[h];
  }
  for (int h = sp->n_heads; h < 4; ++h) action[h] = 0;
  *logp = lp;
  *entropy = en;
}

/* a1: util.action_from_policy (util.py:63-81) -> ActorCriticPolicy.forward, or
 * evaluate_actions when action_in != NULL. */
void orc_policy_forward(const orc_space* sp, const float* params, const void* obs,
                        int64_t obs_stride, int64_t B, uint64_t seed, uint32_t rng_stream,
                        uint32_t tick, uint32_t slot, int64_t idx0, const uint8_t* action_in,
                        uint8_t* action, float* value, float* logp, float* entropy,
                        float* logits) {
  orc_params q;
  split_params(sp, params, &q);
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    orc_acts a;
    const void* row = sp->obs_kind == 0 ? (const void*)((const uint8_t*)obs + b * obs_stride)
                                        : (const void*)((const float*)obs + b * obs_stride);
    forward_one(sp, &q, row, &a);
    uint8_t act[4] = {0, 0, 0, 0};
    uint32_t rnd[4] = {0, 0, 0, 0};
    int sample = action_in == NULL;
    if (sample)
      orc_philox(seed, rng_stream, (uint64_t)(idx0 + b), tick, slot, rnd);
    else
      memcpy(act, action_in + 4 * b, 4);
    float lp, en;
    dist_eval(sp, a.logits, sample, rnd, act, &lp, &en);
    if (action) memcpy(action + 4 * b, act, 4);
    if (value) value[b] = a.value;
    if (logp) logp[b] = lp;
    if (entropy) entropy[b] = en;
    if (logits) memcpy(logits + b * q.L, a.logits, sizeof(float) * q.L);
  }
}

/* ---- AdapPolicyMult (adap/policies.py:134-283, MultModel): per tower (0 = policy, 1 = value)
 *   x = tanh(W0 f + b0);  s_m = tanh(bs_m + sum_k fma(x_k, Ws[m][k])), m = j C + c < 64 C;
 *   y_j = fma(s[j C + C-1], ctx_{C-1}, ... fma(s[j C], ctx_0, x_j));  h2 = tanh(W1 y + b1)
 * on the features WITHOUT the context (MultModel.forward :263-267: x + matmul(x_a.view(B, 64, C), ctx)). */
typedef struct {
  const float *w0[2], *b0[2], *ws[2], *bs[2], *w1[2], *b1[2], *w_act, *b_act, *w_val, *b_val;
  int F, L, C;
} orc_mult_params;
typedef struct {
  float x[2][H], s[2][H * 8], y[2][H], h2[2][H], logits[32 * MAX_HEADS], value;
} orc_mult_acts;

static void split_params_mult(const orc_space* sp, const float* p, int C, orc_mult_params* q) {
  const int F = feat_dim(sp), L = logit_dim(sp);
  q->F = F; q->L = L; q->C = C;
  for (int t = 0; t < 2; ++t) {
    q->w0[t] = p; p += H * F;
    q->b0[t] = p; p += H;
    q->ws[t] = p; p += H * C * H;
    q->bs[t] = p; p += H * C;
    q->w1[t] = p; p += H * H;
    q->b1[t] = p; p += H;
  }
  q->w_act = p; p += L * H;
  q->b_act = p; p += L;
  q->w_val = p; p += H;
  q->b_val = p;
}
int64_t orc_adap_mult_param_count(const orc_space* sp, int32_t C) {
  int64_t F = feat_dim(sp), L = logit_dim(sp);
  return 2 * (H * F + H + (int64_t)H * C * H + H * C + H * H + H) + L * H + L + H + 1;
}

static void forward_one_mult(const orc_space* sp, const orc_mult_params* q, const void* obs_row, const float* ctx,
                             orc_mult_acts* a, int towers) {
  float z[H * 8];
  for (int t = 0; t < towers; ++t) {
    first_layer(sp, obs_row, q->w0[t], q->b0[t], q->F, z);
    for (int j = 0; j < H; ++j) a->x[t][j] = orc_tanhf(z[j]);
    dense(a->x[t], H, q->ws[t], q->bs[t], H * q->C, z);
    for (int m = 0; m < H * q->C; ++m) a->s[t][m] = orc_tanhf(z[m]);
    for (int j = 0; j < H; ++j) {
      float acc = a->x[t][j];
      for (int c = 0; c < q->C; ++c) acc = fmaf(a->s[t][j * q->C + c], ctx[c], acc);
      a->y[t][j] = acc;
    }
    dense(a->y[t], H, q->w1[t], q->b1[t], H, z);
    for (int j = 0; j < H; ++j) a->h2[t][j] = orc_tanhf(z[j]);
  }
  dense(a->h2[0], H, q->w_act, q->b_act, q->L, a->logits);
  a->value = 0.f;
  if (towers > 1) dense(a->h2[1], H, q->w_val, q->b_val, 1, &a->value);
}

void orc_adap_mult_forward(const orc_space* sp, const float* params, int32_t C, const void* obs, int64_t obs_stride,
                           const float* ctx, int64_t ctx_stride, int64_t B, uint64_t seed, uint32_t rng_stream,
                           uint32_t tick, uint32_t slot, int64_t idx0, const uint8_t* action_in, uint8_t* action,
                           float* value, float* logp, float* entropy, float* logits) {
  orc_mult_params q;
  split_params_mult(sp, params, C, &q);
  for (int64_t b = 0; b < B; ++b) {
    orc_mult_acts a;
    const void* row = sp->obs_kind == 0 ? (const void*)((const uint8_t*)obs + b * obs_stride)
                                        : (const void*)((const float*)obs + b * obs_stride);
    forward_one_mult(sp, &q, row, ctx + b * ctx_stride, &a, 2);
    uint8_t act[4] = {0, 0, 0, 0};
    uint32_t rnd[4] = {0, 0, 0, 0};
    int sample = action_in == NULL;
    if (sample)
      orc_philox(seed, rng_stream, (uint64_t)(idx0 + b), tick, slot, rnd);
    else
      memcpy(act, action_in + 4 * b, 4);
    float lp, en;
    dist_eval(sp, a.logits, sample, rnd, act, &lp, &en);
    if (action) memcpy(action + 4 * b, act, 4);
    if (value) value[b] = a.value;
    if (logp) logp[b] = lp;
    if (entropy) entropy[b] = en;
    if (logits) memcpy(logits + b * q.L, a.logits, sizeof(float) * q.L);
  }
}

/* AdapPolicy.forward / evaluate_actions (adap/policies.py:86-131): orc_policy_forward with C context
 * inputs per sample (ctx_stride = 0: one context for the whole batch, `self.context.repeat`). */
void orc_adap_forward(const orc_space* sp, const float* params, int32_t C, const void* obs, int64_t obs_stride,
                      const float* ctx, int64_t ctx_stride, int64_t B, uint64_t seed, uint32_t rng_stream,
                      uint32_t tick, uint32_t slot, int64_t idx0, const uint8_t* action_in, uint8_t* action,
                      float* value, float* logp, float* entropy, float* logits) {
  orc_params q;
  split_params_ctx(sp, params, C, &q);
  for (int64_t b = 0; b < B; ++b) {
    orc_acts a;
    const void* row = sp->obs_kind == 0 ? (const void*)((const uint8_t*)obs + b * obs_stride)
                                        : (const void*)((const float*)obs + b * obs_stride);
    forward_one_ctx(sp, &q, row, ctx + b * ctx_stride, &a);
    uint8_t act[4] = {0, 0, 0, 0};
    uint32_t rnd[4] = {0, 0, 0, 0};
    int sample = action_in == NULL;
    if (sample)
      orc_philox(seed, rng_stream, (uint64_t)(idx0 + b), tick, slot, rnd);
    else
      memcpy(act, action_in + 4 * b, 4);
    float lp, en;
    dist_eval(sp, a.logits, sample, rnd, act, &lp, &en);
    if (action) memcpy(action + 4 * b, act, 4);
    if (value) value[b] = a.value;
    if (logp) logp[b] = lp;
    if (entropy) entropy[b] = en;
    if (logits) memcpy(logits + b * q.L, a.logits, sizeof(float) * q.L);
  }
}

#include "pth_oracle_overcooked.inc"
#include "pth_oracle_rollout.inc"
#include "pth_oracle_update.inc"
#include "pth_oracle_modular.inc"
