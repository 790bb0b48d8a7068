// pth_rollout.cu — the T-tick rollout megakernel: ego forward + sample, env
// step, partner forward + sample, reward routing, auto-reset and all rollout
// buffer writes for 128 env instances per CTA, without returning to the host.
//
// Replaces the Python loop  SB3 OnPolicyAlgorithm.collect_rollouts (restated
// pantheonrl/algos/adap/adap_learn.py:415-471) -> MultiAgentEnv.step / reset
// (pantheonrl/common/multiagentenv.py:172-243) -> OnPolicyAgent.get_action /
// update (pantheonrl/common/agents.py:111-203) -> RPSEnv / LiarEnv rules.
//
// Envs are independent and both policies are frozen during a rollout, so a CTA
// owns its 128 envs for the whole horizon: weights of both agents live in
// shared memory, game state and routing latches live in registers, and every
// tick runs up to three CTA-wide batched forwards (ego; partner reply; partner
// opening move after an auto-reset) with lane masks.  Event order and RNG slots
// follow oracle/pth_oracle_rollout.inc exactly.
#include <stdlib.h>

#include "pth_games.cuh"
#include "pth_mlp.cuh"
#include "pth_overcooked.cuh"

using namespace pthmlp;

namespace {

// RB = envs per CTA: 128 when there are enough envs to give every SM a tile,
// 32 otherwise (N = 4096 -> 128 CTAs instead of 32).
template <int RB>
struct RollSmem {
  SmemPolicy pol_ego;
  SmemPolicy pol_alt;
  float A[HID * (RB + 4)];
  float Bf[HID * (RB + 4)];
  float Lg[MAXL * (RB + 4)];
  uint32_t obs[RB * 8];
  float red[32];
};

struct RollParams {
  SpaceDev sp;
  Layout lo;
  pth_rollout_args a;
};

struct FwdOut {
  uint32_t action;
  float value, logp;
};

// One CTA-wide forward over the observations currently staged in shared memory:
// BOX = false: obs_s = [RB][32] bytes (one-hot slots); BOX = true: obs_s = feature-major
// fp32 tile X[k][b] (row stride RB + 4).
template <int RB, bool BOX>
__device__ __forceinline__ FwdOut cta_forward_t(const RollParams& p, const float* __restrict__ params,
                                                const SmemPolicy& pol, float* A, float* Bf, float* Lg,
                                                const void* obs_s, int tid, bool with_policy, pth_u4 rnd) {
  FwdOut o;
  o.action = 0;
  o.logp = 0.f;
  __syncthreads();  // obs written by all lanes; previous users of A/Bf/Lg are done
  const bool lane = tid < RB;  // threads [RB, NT) only help in the tiled layers
  auto first = [&](const float* W, const float* bias) {
    if constexpr (BOX)
      first_layer_box<false, NT, RB>(p.sp.F, reinterpret_cast<const float*>(obs_s), W, bias, A, tid);
    else
      first_layer_onehot<false, NT, RB>(p.sp, reinterpret_cast<const uint8_t*>(obs_s), W, bias, A, tid);
  };
  if (with_policy) {
    first(params + p.lo.w_pi0, pol.b_pi0);
    __syncthreads();
    dense64<true, NT, RB>(A, pol.w_pi1, pol.b_pi1, Bf, tid);
    __syncthreads();
    if (lane) action_head<RB>(Bf, pol, p.sp.L, Lg, tid);
  }
  first(params + p.lo.w_vf0, pol.b_vf0);
  __syncthreads();
  dense64<true, NT, RB>(A, pol.w_vf1, pol.b_vf1, Bf, tid);
  __syncthreads();
  o.value = 0.f;
  if (!lane) return o;
  o.value = value_head<RB>(Bf, pol, tid);
  if (with_policy) {
    DistOut d = dist_eval<RB>(p.sp, Lg, tid, true, rnd, 0u);
    o.action = d.action;
    o.logp = d.logp;
  }
  return o;
}

template <int RB>
__device__ __forceinline__ FwdOut cta_forward(const RollParams& p, const float* __restrict__ params,
                                              const SmemPolicy& pol, RollSmem<RB>& sm, int tid,
                                              bool with_policy, pth_u4 rnd) {
  return cta_forward_t<RB, false>(p, params, pol, sm.A, sm.Bf, sm.Lg, sm.obs, tid, with_policy, rnd);
}

struct EnvRegs {
  LiarRegs liar;
  float total_ego, total_alt;
  float ego_last_start, alt_last_done;
  float alt_pending;   // reward accumulated on the partner's latest row
  float alt_row_start; // episode_start stored with the partner's latest row
  int alt_count;
  uint32_t flags;      // bit0 ego_moved, bit1 should_update, bit2 open partner row carried
  float st_eps, st_rew, st_steps, st_alt;
};

__device__ __forceinline__ void store_obs_row(uint8_t* base, int64_t row, const uint32_t (&w)[8]) {
  uint4* q = reinterpret_cast<uint4*>(base + row * 32);
  q[0] = make_uint4(w[0], w[1], w[2], w[3]);
  q[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// OnPolicyAgent.update (agents.py:196-198)
__device__ __forceinline__ void alt_update(EnvRegs& e, float reward, bool done) {
  e.alt_last_done = done ? 1.f : 0.f;
  if (e.alt_count > 0) e.alt_pending = e.alt_pending + reward;
}

__device__ __forceinline__ void alt_flush(const RollParams& p, const EnvRegs& e, int64_t n) {
  if (p.a.partner_records && e.alt_count > 0)
    p.a.alt.d_rewards[(int64_t)(e.alt_count - 1) * p.a.N + n] = e.alt_pending;
}

// partner.get_action bookkeeping after the batched forward: record the row
// (agents.py:172-179), then the first-move hand-off (multiagentenv.py:158-160).
__device__ __forceinline__ void alt_record(const RollParams& p, EnvRegs& e, int64_t n,
                                           const uint32_t (&obs)[8], const FwdOut& f) {
  if (p.a.partner_records && e.alt_count < p.a.alt.Tcap) {
    alt_flush(p, e, n);
    const int64_t o = (int64_t)e.alt_count * p.a.N + n;
    store_obs_row(p.a.alt.d_obs, o, obs);
    *reinterpret_cast<uint32_t*>(p.a.alt.d_actions + 4 * o) = f.action;
    p.a.alt.d_values[o] = f.value;
    p.a.alt.d_logp[o] = f.logp;
    p.a.alt.d_episode_starts[o] = e.alt_last_done;
    e.alt_row_start = e.alt_last_done;
    e.alt_count += 1;
    e.alt_pending = 0.f;
  }
  e.st_alt += 1.f;
  if (!(e.flags & 2u)) alt_update(e, e.total_alt, false);
  e.flags |= 2u;
}

// MultiAgentEnv._update_players (multiagentenv.py:163-170)
__device__ __forceinline__ void update_players(EnvRegs& e, float r_ego, float r_alt, bool done) {
  if (e.flags & 2u) alt_update(e, r_alt, done);
  e.total_ego = e.total_ego + r_ego;
  e.total_alt = e.total_alt + r_alt;
}

template <int ENV, int RB>
__global__ void __launch_bounds__(NT) rollout_kernel(const __grid_constant__ RollParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RollSmem<RB>& sm = *reinterpret_cast<RollSmem<RB>*>(smem_raw);
  const int tid = threadIdx.x;
  const bool lane = tid < RB;  // one env per thread in [0, RB); the rest help with the MLP tiles
  const int64_t n = (int64_t)blockIdx.x * RB + (lane ? tid : 0);
  const int64_t N = p.a.N;
  const bool valid = lane && n < N;
  const uint64_t genv = (uint64_t)(p.a.env0 + n);
  const bool selfplay = p.a.d_alt_params == p.a.d_ego_params;
  const float* ego_w = p.a.d_ego_params;
  const float* alt_w = p.a.d_alt_params;

  load_policy(sm.pol_ego, ego_w, p.lo, p.sp.L, tid, NT);
  if (!selfplay) load_policy(sm.pol_alt, alt_w, p.lo, p.sp.L, tid, NT);
  const SmemPolicy& pol_alt = selfplay ? sm.pol_ego : sm.pol_alt;
  if (lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm.obs[tid * 8 + i] = 0u;
  }

  EnvRegs e;
  memset(&e, 0, sizeof(e));
  const pth_env_carry& cr = p.a.carry;
  if (valid) {
    if (p.a.first_rollout) {
      e.ego_last_start = 1.f;  // SB3 _setup_learn: _last_episode_starts = ones
      e.alt_last_done = 1.f;   // agents.py:97
    } else {
      e.ego_last_start = cr.d_ego_last_start[n];
      e.alt_last_done = cr.d_alt_last_done[n];
      e.total_ego = cr.d_total_rew[n];
      e.total_alt = cr.d_total_rew[N + n];
      e.flags = cr.d_flags[n];
      if (ENV == PTH_ENV_LIAR) {
        liar_load(reinterpret_cast<const pth_liar_state*>(cr.d_game_state) + n, e.liar);
        if (e.flags & 4u) {
          // The partner row that was still waiting for the ego's next move when the previous rollout
          // ended (it sits at row count[n] of the partner's buffer) becomes row 0 of this rollout.
          if (p.a.partner_records) {
            const int64_t src = (int64_t)p.a.alt.d_count[n] * N + n;
            const uint4* q = reinterpret_cast<const uint4*>(p.a.alt.d_obs + src * 32);
            const uint4 o0 = q[0], o1 = q[1];
            uint4* d = reinterpret_cast<uint4*>(p.a.alt.d_obs + n * 32);
            d[0] = o0;
            d[1] = o1;
            *reinterpret_cast<uint32_t*>(p.a.alt.d_actions + 4 * n) =
                *reinterpret_cast<const uint32_t*>(p.a.alt.d_actions + 4 * src);
            p.a.alt.d_values[n] = p.a.alt.d_values[src];
            p.a.alt.d_logp[n] = p.a.alt.d_logp[src];
            e.alt_row_start = p.a.alt.d_episode_starts[src];
            p.a.alt.d_episode_starts[n] = e.alt_row_start;
            e.alt_pending = p.a.alt.d_rewards[src];
            e.alt_count = 1;
          }
          e.flags &= ~4u;
        }
      }
    }
  }

  uint32_t w[8];
  // ---- initial reset (MultiAgentEnv.reset from SB3 _setup_learn)
  if (ENV == PTH_ENV_LIAR && p.a.first_rollout) {
    bool ego_first = true;
    if (valid) ego_first = liar_reset(e.liar, p.a.seed, genv, p.a.tick0, 8u, p.a.probegostart);
    const bool act_c = valid && !ego_first;
    if (__syncthreads_or(act_c)) {
      if (act_c) {
        liar_obs(e.liar, 1, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.obs[tid * 8 + i] = w[i];
      }
      pth_u4 rnd = pth_philox(p.a.seed, PTH_STREAM_ALT, genv, p.a.tick0, 2u);
      FwdOut f = cta_forward<RB>(p, alt_w, pol_alt, sm, tid, true, rnd);
      if (act_c) {
        alt_record(p, e, n, w, f);
        float re, ra;
        liar_step(e.liar, 1, (int)(f.action & 0xffu), (int)((f.action >> 8) & 0xffu), re, ra);
        update_players(e, re, ra, false);
      }
    }
  }

  for (int64_t t = 0; t < p.a.T; ++t) {
    const uint32_t g = p.a.tick0 + (uint32_t)t;
    const int64_t o = t * N + n;
    // ================= ego decision
    if (ENV == PTH_ENV_LIAR) {
      liar_obs(e.liar, 0, w);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = 0u;
    }
    __syncthreads();  // previous forward finished reading sm.obs
    if (lane) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm.obs[tid * 8 + i] = w[i];
    }
    FwdOut fe = cta_forward<RB>(p, ego_w, sm.pol_ego, sm, tid, true,
                            pth_philox(p.a.seed, PTH_STREAM_EGO, genv, g, 0u));
    if (valid) {
      store_obs_row(p.a.ego.d_obs, o, w);
      *reinterpret_cast<uint32_t*>(p.a.ego.d_actions + 4 * o) = fe.action;
      p.a.ego.d_values[o] = fe.value;
      p.a.ego.d_logp[o] = fe.logp;
      p.a.ego.d_episode_starts[o] = e.ego_last_start;
    }
    float ego_rew = 0.f;
    bool done = false;
    if (ENV == PTH_ENV_RPS) {
      // ================= SimultaneousEnv: partner acts on the same tick
      __syncthreads();
      FwdOut fa = cta_forward<RB>(p, alt_w, pol_alt, sm, tid, true,
                              pth_philox(p.a.seed, PTH_STREAM_ALT, genv, g, 0u));
      if (valid) {
        alt_record(p, e, n, w, fa);
        float re, ra;
        pth_rps_outcome((int)(fe.action & 0xffu), (int)(fa.action & 0xffu), re, ra);
        done = true;
        update_players(e, re, ra, done);
        ego_rew = ego_rew + ((e.flags & 1u) ? re : e.total_ego);
        e.flags |= 1u;
      }
    } else {
      // ================= TurnBasedEnv: ego_step, then the partner's reply
      float re = 0.f, ra = 0.f;
      if (valid) {
        done = liar_step(e.liar, 0, (int)(fe.action & 0xffu), (int)((fe.action >> 8) & 0xffu), re, ra);
        update_players(e, re, ra, done);
        ego_rew = ego_rew + ((e.flags & 1u) ? re : e.total_ego);
        e.flags |= 1u;
      }
      const bool act_b = valid && !done;
      if (__syncthreads_or(act_b)) {
        if (act_b) {
          liar_obs(e.liar, 1, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) sm.obs[tid * 8 + i] = w[i];
        }
        FwdOut fa = cta_forward<RB>(p, alt_w, pol_alt, sm, tid, true,
                                pth_philox(p.a.seed, PTH_STREAM_ALT, genv, g, 0u));
        if (act_b) {
          alt_record(p, e, n, w, fa);
          done = liar_step(e.liar, 1, (int)(fa.action & 0xffu), (int)((fa.action >> 8) & 0xffu), re, ra);
          update_players(e, re, ra, done);
          ego_rew = ego_rew + re;
        }
      }
    }
    if (valid) {
      p.a.ego.d_rewards[o] = ego_rew;
      e.ego_last_start = done ? 1.f : 0.f;
      e.st_steps += 1.f;
      if (done) {
        e.st_eps += 1.f;
        e.st_rew += e.total_ego;
      }
    }
    // ================= DummyVecEnv auto-reset -> MultiAgentEnv.reset
    if (ENV == PTH_ENV_RPS) {
      if (valid && done) {
        e.flags = 0u;
        e.total_ego = 0.f;
        e.total_alt = 0.f;
      }
    } else {
      bool ego_first = true;
      if (valid && done) {
        ego_first = liar_reset(e.liar, p.a.seed, genv, g, 0u, p.a.probegostart);
        e.flags = 0u;
        e.total_ego = 0.f;
        e.total_alt = 0.f;
      }
      const bool act_c = valid && done && !ego_first;
      if (__syncthreads_or(act_c)) {
        if (act_c) {
          liar_obs(e.liar, 1, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) sm.obs[tid * 8 + i] = w[i];
        }
        FwdOut fa = cta_forward<RB>(p, alt_w, pol_alt, sm, tid, true,
                                pth_philox(p.a.seed, PTH_STREAM_ALT, genv, g, 1u));
        if (act_c) {
          alt_record(p, e, n, w, fa);
          float re, ra;
          liar_step(e.liar, 1, (int)(fa.action & 0xffu), (int)((fa.action >> 8) & 0xffu), re, ra);
          update_players(e, re, ra, false);
        }
      }
    }
  }

  // ---- bootstrap value of the ego's next observation (SB3 predict_values)
  if (ENV == PTH_ENV_LIAR) {
    liar_obs(e.liar, 0, w);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 0u;
  }
  __syncthreads();
  if (lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sm.obs[tid * 8 + i] = w[i];
  }
  pth_u4 zero = {0, 0, 0, 0};
  FwdOut fl = cta_forward<RB>(p, ego_w, sm.pol_ego, sm, tid, false, zero);

  if (valid) {
    alt_flush(p, e, n);
    float boot_done = e.alt_last_done;
    if (ENV == PTH_ENV_LIAR && p.a.partner_records && (e.flags & 2u) && e.alt_count > 0) {
      // turn-based, the partner has moved in the running episode: its latest row still waits for the
      // ego's next move.  It is left out of this rollout's batch and carried.
      e.alt_count -= 1;
      boot_done = e.alt_row_start;
      e.flags |= 4u;
    }
    if (cr.d_alt_boot_done) cr.d_alt_boot_done[n] = boot_done;
    cr.d_ego_last_value[n] = fl.value;
    cr.d_ego_last_done[n] = e.ego_last_start;
    cr.d_ego_last_start[n] = e.ego_last_start;
    cr.d_alt_last_done[n] = e.alt_last_done;
    cr.d_total_rew[n] = e.total_ego;
    cr.d_total_rew[N + n] = e.total_alt;
    cr.d_flags[n] = (uint8_t)e.flags;
    if (ENV == PTH_ENV_LIAR)
      liar_store(reinterpret_cast<pth_liar_state*>(cr.d_game_state) + n, e.liar);
    if (p.a.alt.d_count) p.a.alt.d_count[n] = e.alt_count;
  }
  // ---- episode statistics (integer-valued floats: exact in any order)
  if (cr.d_ep_stats) {
    float v[4] = {e.st_eps, e.st_rew, e.st_steps, e.st_alt};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x = v[i];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      if ((tid & 31) == 0) atomicAdd(cr.d_ep_stats + i, x);
    }
  }
}

// ---------------------------------------------------------------- Overcooked
// SimultaneousEnv with per-agent Box observations (overcookedgym/overcooked.py:10-98 behind
// multiagentenv.py:149-243, 395-409).  Per tick: both agents' feature rows are derived once
// from the env's registers into two shared-memory tiles (Xe, Xa), stored to the rollout
// buffers with coalesced 256-byte rows, ego forward on Xe, partner forward on Xa, one joint
// env step, reward routing, horizon auto-reset.  No env randomness (standard start state).
template <int RB>
struct RollSmemOC {
  SmemPolicy pol_ego;
  SmemPolicy pol_alt;
  float A[HID * (RB + 4)];
  float Bf[HID * (RB + 4)];
  float Lg[MAXL * (RB + 4)];
  float Xe[HID * (RB + 4)];
  float Xa[HID * (RB + 4)];
  pth_overcooked_layout lay;
};

// rows of the tile X (feature-major) -> obs rows [row0 + b] of 64 floats, 16 threads per row
template <int RB>
__device__ __forceinline__ void store_box_rows(const float* X, float* dst_rows, int64_t row0, int64_t n_valid,
                                               int tid) {
  constexpr int LDX = RB + 4;
  for (int i = tid; i < RB * 16; i += NT) {
    const int b = i >> 4, k4 = (i & 15) * 4;
    if (b < n_valid) {
      const float4 v = make_float4(X[(k4 + 0) * LDX + b], X[(k4 + 1) * LDX + b], X[(k4 + 2) * LDX + b],
                                   X[(k4 + 3) * LDX + b]);
      __stcs(reinterpret_cast<float4*>(dst_rows + (row0 + b) * PTH_OC_ROW + k4), v);
    }
  }
}

template <int RB>
__global__ void __launch_bounds__(NT) rollout_overcooked_kernel(const __grid_constant__ RollParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RollSmemOC<RB>& sm = *reinterpret_cast<RollSmemOC<RB>*>(smem_raw);
  constexpr int LDX = RB + 4;
  const int tid = threadIdx.x;
  const bool lane = tid < RB;
  const int64_t n0 = (int64_t)blockIdx.x * RB;
  const int64_t n = n0 + (lane ? tid : 0);
  const int64_t N = p.a.N;
  const bool valid = lane && n < N;
  const int64_t n_valid = (N - n0 < RB) ? (N - n0) : RB;
  const uint64_t genv = (uint64_t)(p.a.env0 + n);
  const bool selfplay = p.a.d_alt_params == p.a.d_ego_params;
  const float* ego_w = p.a.d_ego_params;
  const float* alt_w = p.a.d_alt_params;

  load_policy(sm.pol_ego, ego_w, p.lo, p.sp.L, tid, NT);
  if (!selfplay) load_policy(sm.pol_alt, alt_w, p.lo, p.sp.L, tid, NT);
  const SmemPolicy& pol_alt = selfplay ? sm.pol_ego : sm.pol_alt;
  {
    const uint2* src = reinterpret_cast<const uint2*>(p.a.d_layout);
    uint2* dst = reinterpret_cast<uint2*>(&sm.lay);
    for (int i = tid; i < (int)(sizeof(pth_overcooked_layout) / 8); i += NT) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < HID * LDX; i += NT) {
    sm.Xe[i] = 0.f;
    sm.Xa[i] = 0.f;
  }
  __syncthreads();
  const pth_overcooked_layout& L = sm.lay;

  EnvRegs e;
  memset(&e, 0, sizeof(e));
  OcRegs g;
  oc_reset(L, g);
  const pth_env_carry& cr = p.a.carry;
  pth_overcooked_state* gstate = reinterpret_cast<pth_overcooked_state*>(cr.d_game_state);
  if (valid) {
    if (p.a.first_rollout) {
      e.ego_last_start = 1.f;  // SB3 _setup_learn: _last_episode_starts = ones
      e.alt_last_done = 1.f;   // agents.py:97
    } else {
      e.ego_last_start = cr.d_ego_last_start[n];
      e.alt_last_done = cr.d_alt_last_done[n];
      e.total_ego = cr.d_total_rew[n];
      e.total_alt = cr.d_total_rew[N + n];
      e.flags = cr.d_flags[n];
      oc_load(gstate + n, g);
    }
  }
  float* ego_rows = reinterpret_cast<float*>(p.a.ego.d_obs);
  float* alt_rows = reinterpret_cast<float*>(p.a.alt.d_obs);

  for (int64_t t = 0; t < p.a.T; ++t) {
    const uint32_t gt = p.a.tick0 + (uint32_t)t;
    const int64_t o = t * N + n;
    __syncthreads();  // the previous tick's forwards are done with Xe / Xa
    if (lane) oc_write_obs(L, g, sm.Xe, sm.Xa, LDX, tid);
    __syncthreads();
    store_box_rows<RB>(sm.Xe, ego_rows, t * N + n0, n_valid, tid);
    const bool alt_rec = p.a.partner_records && t < p.a.alt.Tcap;  // simultaneous: one partner row per tick
    if (alt_rec) store_box_rows<RB>(sm.Xa, alt_rows, t * N + n0, n_valid, tid);
    // ================= ego decision, partner decision (same tick)
    FwdOut fe = cta_forward_t<RB, true>(p, ego_w, sm.pol_ego, sm.A, sm.Bf, sm.Lg, sm.Xe, tid, true,
                                        pth_philox(p.a.seed, PTH_STREAM_EGO, genv, gt, 0u));
    FwdOut fa = cta_forward_t<RB, true>(p, alt_w, pol_alt, sm.A, sm.Bf, sm.Lg, sm.Xa, tid, true,
                                        pth_philox(p.a.seed, PTH_STREAM_ALT, genv, gt, 0u));
    if (valid) {
      *reinterpret_cast<uint32_t*>(p.a.ego.d_actions + 4 * o) = fe.action;
      p.a.ego.d_values[o] = fe.value;
      p.a.ego.d_logp[o] = fe.logp;
      p.a.ego.d_episode_starts[o] = e.ego_last_start;
      // partner.get_action bookkeeping (agents.py:172-179) + first-move hand-off (multiagentenv.py:158-160)
      if (alt_rec) {
        alt_flush(p, e, n);
        const int64_t oa = (int64_t)e.alt_count * N + n;
        *reinterpret_cast<uint32_t*>(p.a.alt.d_actions + 4 * oa) = fa.action;
        p.a.alt.d_values[oa] = fa.value;
        p.a.alt.d_logp[oa] = fa.logp;
        p.a.alt.d_episode_starts[oa] = e.alt_last_done;
        e.alt_count += 1;
        e.alt_pending = 0.f;
      }
      e.st_alt += 1.f;
      if (!(e.flags & 2u)) alt_update(e, e.total_alt, false);
      e.flags |= 2u;
      // OvercookedMultiEnv.multi_step: joint action in player order, one reward for both
      const int ea = (int)(fe.action & 0xffu), aa = (int)(fa.action & 0xffu);
      float rew;
      const bool done = L.ego_agent_idx == 0 ? oc_step(L, g, ea, aa, rew) : oc_step(L, g, aa, ea, rew);
      update_players(e, rew, rew, done);
      const float ego_rew = (e.flags & 1u) ? rew : e.total_ego;
      e.flags |= 1u;
      p.a.ego.d_rewards[o] = ego_rew;
      e.ego_last_start = done ? 1.f : 0.f;
      e.st_steps += 1.f;
      if (done) {  // DummyVecEnv auto-reset -> MultiAgentEnv.reset -> multi_reset
        e.st_eps += 1.f;
        e.st_rew += e.total_ego;
        e.flags = 0u;
        e.total_ego = 0.f;
        e.total_alt = 0.f;
        oc_reset(L, g);
      }
    }
  }

  // ---- bootstrap value of the ego's next observation (SB3 predict_values)
  __syncthreads();
  if (lane) oc_write_obs(L, g, sm.Xe, sm.Xa, LDX, tid);
  pth_u4 zero = {0, 0, 0, 0};
  FwdOut fl = cta_forward_t<RB, true>(p, ego_w, sm.pol_ego, sm.A, sm.Bf, sm.Lg, sm.Xe, tid, false, zero);
  if (valid) {
    alt_flush(p, e, n);
    if (cr.d_alt_boot_done) cr.d_alt_boot_done[n] = e.alt_last_done;  // simultaneous: no open rows
    cr.d_ego_last_value[n] = fl.value;
    cr.d_ego_last_done[n] = e.ego_last_start;
    cr.d_ego_last_start[n] = e.ego_last_start;
    cr.d_alt_last_done[n] = e.alt_last_done;
    cr.d_total_rew[n] = e.total_ego;
    cr.d_total_rew[N + n] = e.total_alt;
    cr.d_flags[n] = (uint8_t)e.flags;
    oc_store(gstate + n, g);
    if (p.a.alt.d_count) p.a.alt.d_count[n] = e.alt_count;
  }
  if (cr.d_ep_stats) {
    float v[4] = {e.st_eps, e.st_rew, e.st_steps, e.st_alt};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x = v[i];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      if ((tid & 31) == 0) atomicAdd(cr.d_ep_stats + i, x);
    }
  }
}

}  // namespace

extern "C" int pth_rollout_run(pth_ctx* ctx, const pth_rollout_args* a, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr && a != nullptr, "NULL ctx/args");
  PTH_CHECK_ARG(a->space && a->d_ego_params && a->d_alt_params, "NULL space/params");
  PTH_CHECK_ARG(a->env_kind == PTH_ENV_RPS || a->env_kind == PTH_ENV_LIAR ||
                    a->env_kind == PTH_ENV_OVERCOOKED,
                "unknown env_kind");
  PTH_CHECK_ARG(a->N >= 0 && a->T >= 0, "negative size");
  PTH_CHECK_ARG(a->ego.d_obs && a->ego.d_actions && a->ego.d_rewards && a->ego.d_values &&
                    a->ego.d_logp && a->ego.d_episode_starts,
                "NULL ego buffer");
  PTH_CHECK_ARG(a->ego.Tcap >= a->T, "ego buffer shorter than T");
  if (a->partner_records) {
    PTH_CHECK_ARG(a->alt.d_obs && a->alt.d_actions && a->alt.d_rewards && a->alt.d_values &&
                      a->alt.d_logp && a->alt.d_episode_starts && a->alt.d_count,
                  "NULL partner buffer");
    PTH_CHECK_ARG(a->alt.Tcap >= (a->env_kind == PTH_ENV_LIAR ? 2 * a->T + 1 : a->T),
                  "partner buffer must hold 2*T + 1 rows (T for simultaneous envs)");
  }
  const pth_env_carry& c = a->carry;
  PTH_CHECK_ARG(c.d_ego_last_start && c.d_alt_last_done && c.d_total_rew && c.d_flags &&
                    c.d_ego_last_value && c.d_ego_last_done,
                "NULL carry array");
  PTH_CHECK_ARG(a->env_kind == PTH_ENV_RPS || c.d_game_state, "NULL game state");
  PTH_CHECK_ARG(((uintptr_t)a->ego.d_obs % 16) == 0 &&
                    (!a->partner_records || ((uintptr_t)a->alt.d_obs % 16) == 0) &&
                    ((uintptr_t)c.d_game_state % 16) == 0 && ((uintptr_t)a->d_ego_params % 16) == 0 &&
                    ((uintptr_t)a->d_alt_params % 16) == 0,
                "obs/state/params must be 16-byte aligned");
  if (a->N == 0) return PTH_OK;
  RollParams p;
  if (a->env_kind == PTH_ENV_OVERCOOKED) {
    PTH_CHECK_ARG(a->d_layout != nullptr && ((uintptr_t)a->d_layout % 8) == 0, "NULL / misaligned d_layout");
    if (fill_space(a->space, &p.sp) != 0 || p.sp.obs_kind != PTH_OBS_BOX || p.sp.obs_len != PTH_OC_OBS ||
        p.sp.n_heads != 1 || p.sp.head_n[0] != 6) {
      pth_set_error("pth_rollout_run: Overcooked needs Box(62) observations and Discrete(6) actions");
      return PTH_EINVAL;
    }
    p.lo = make_layout(p.sp.F, p.sp.L);
    p.a = *a;
    // 16-env tiles while 32-env tiles would leave SMs idle (1024 envs: 64 CTAs instead of 32)
    if (pth_ceil_div(a->N, 32) < ctx->sm_count && !getenv("PTH_ROLLOUT_RB32")) {
      constexpr int RBV = 16;
      const size_t smem = sizeof(RollSmemOC<RBV>);
      PTH_CUDA(cudaFuncSetAttribute(rollout_overcooked_kernel<RBV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      rollout_overcooked_kernel<RBV><<<pth_ceil_div(a->N, RBV), NT, smem, (cudaStream_t)stream>>>(p);
    } else {
      constexpr int RBV = 32;
      const size_t smem = sizeof(RollSmemOC<RBV>);
      PTH_CUDA(cudaFuncSetAttribute(rollout_overcooked_kernel<RBV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      rollout_overcooked_kernel<RBV><<<pth_ceil_div(a->N, RBV), NT, smem, (cudaStream_t)stream>>>(p);
    }
    PTH_LAUNCH_CHECK();
    return PTH_OK;
  }
  if (fill_space(a->space, &p.sp) != 0 || p.sp.obs_kind != PTH_OBS_ONEHOT) {
    pth_set_error("pth_rollout_run: RPS / Liar's Dice need a one-hot observation space");
    return PTH_ENOSUP;
  }
  const int want_len = a->env_kind == PTH_ENV_LIAR ? PTH_LIAR_OBS_LEN : 1;
  const int want_heads = a->env_kind == PTH_ENV_LIAR ? 2 : 1;
  if (p.sp.obs_len != want_len || p.sp.n_heads != want_heads) {
    pth_set_error("pth_rollout_run: space does not match env_kind %d", a->env_kind);
    return PTH_EINVAL;
  }
  p.lo = make_layout(p.sp.F, p.sp.L);
  p.a = *a;
  cudaStream_t st = (cudaStream_t)stream;
  // 128-env tiles once every SM gets one; 32-env tiles below that (4x more CTAs)
  const bool small = pth_ceil_div(a->N, 128) < ctx->sm_count;
#define PTH_ROLL_LAUNCH(ENVK, RBV)                                                              \
  do {                                                                                          \
    const size_t smem = sizeof(RollSmem<RBV>);                                                  \
    PTH_CUDA(cudaFuncSetAttribute(rollout_kernel<ENVK, RBV>,                                    \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    rollout_kernel<ENVK, RBV><<<pth_ceil_div(a->N, RBV), NT, smem, st>>>(p);                    \
  } while (0)
  // 16-env tiles (101 KB of shared memory: two CTAs per SM, 16 warps to hide the row-gather latency)
  // while 32-env tiles would leave SMs with a single CTA
  const bool tiny = pth_ceil_div(a->N, 32) < 2 * ctx->sm_count && !getenv("PTH_ROLLOUT_RB32");
  if (a->env_kind == PTH_ENV_RPS) {
    if (tiny) PTH_ROLL_LAUNCH(PTH_ENV_RPS, 16); else if (small) PTH_ROLL_LAUNCH(PTH_ENV_RPS, 32); else PTH_ROLL_LAUNCH(PTH_ENV_RPS, 128);
  } else {
    if (tiny) PTH_ROLL_LAUNCH(PTH_ENV_LIAR, 16); else if (small) PTH_ROLL_LAUNCH(PTH_ENV_LIAR, 32); else PTH_ROLL_LAUNCH(PTH_ENV_LIAR, 128);
  }
#undef PTH_ROLL_LAUNCH
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}
