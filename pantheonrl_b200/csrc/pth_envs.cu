// pth_envs.cu — standalone vectorised env step / reset kernels (one thread per
// env).  Used by the Python facade when the caller drives envs step by step and
// by the parity tests that replay the reference's own env traces.
#include "pth_games.cuh"
#include "pth_overcooked.cuh"

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

// ---------------------------------------------------------------- Overcooked
// obs out: [N][2][PTH_OC_ROW] fp32 (ego row, partner row).  The layout (5.6 KB) is
// staged in shared memory once per CTA; each thread owns one env.
constexpr int OC_TB = 128;

__device__ __forceinline__ void oc_stage_layout(pth_overcooked_layout* dst, const pth_overcooked_layout* src,
                                                int tid, int nthreads) {
  static_assert(sizeof(pth_overcooked_layout) % 8 == 0, "layout is copied in 8-byte words");
  const uint2* s = reinterpret_cast<const uint2*>(src);
  uint2* d = reinterpret_cast<uint2*>(dst);
  for (int i = tid; i < (int)(sizeof(pth_overcooked_layout) / 8); i += nthreads) d[i] = __ldg(s + i);
}

__device__ __forceinline__ void oc_emit_obs(const pth_overcooked_layout& L, const OcRegs& s, float* obs_n) {
  // thread-private feature-major scratch in registers/local: two rows of 64
  float xe[PTH_OC_ROW], xa[PTH_OC_ROW];
  oc_write_obs(L, s, xe, xa, 1, 0);
#pragma unroll
  for (int k = 0; k < PTH_OC_ROW; k += 4) {
    reinterpret_cast<float4*>(obs_n)[k / 4] = make_float4(xe[k], xe[k + 1], xe[k + 2], xe[k + 3]);
    reinterpret_cast<float4*>(obs_n + PTH_OC_ROW)[k / 4] = make_float4(xa[k], xa[k + 1], xa[k + 2], xa[k + 3]);
  }
}

__global__ void __launch_bounds__(OC_TB) overcooked_reset_kernel(const pth_overcooked_layout* __restrict__ lay,
                                                                 pth_overcooked_state* __restrict__ state,
                                                                 float* __restrict__ obs, int64_t N) {
  __shared__ __align__(16) pth_overcooked_layout L;
  oc_stage_layout(&L, lay, threadIdx.x, OC_TB);
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * OC_TB + threadIdx.x;
  if (n >= N) return;
  OcRegs s;
  oc_reset(L, s);
  oc_store(state + n, s);
  if (obs) oc_emit_obs(L, s, obs + n * 2 * PTH_OC_ROW);
}

__global__ void __launch_bounds__(OC_TB) overcooked_step_kernel(
    const pth_overcooked_layout* __restrict__ lay, pth_overcooked_state* __restrict__ state,
    const uint8_t* __restrict__ ego_a, const uint8_t* __restrict__ alt_a, float* __restrict__ obs,
    float* __restrict__ reward, uint8_t* __restrict__ done, int64_t N) {
  __shared__ __align__(16) pth_overcooked_layout L;
  oc_stage_layout(&L, lay, threadIdx.x, OC_TB);
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * OC_TB + threadIdx.x;
  if (n >= N) return;
  OcRegs s;
  oc_load(state + n, s);
  const int ea = ego_a[n], aa = alt_a[n];
  float r;
  const bool d = L.ego_agent_idx == 0 ? oc_step(L, s, ea, aa, r) : oc_step(L, s, aa, ea, r);
  oc_store(state + n, s);
  if (obs) oc_emit_obs(L, s, obs + n * 2 * PTH_OC_ROW);
  reward[n] = r;
  done[n] = d ? 1 : 0;
}

}  // namespace

// HOST: derived tables of a layout.  The MotionPlanner graph (planners.py:197-222) has one
// node per (floor cell, orientation) and an edge per action; plan lengths are BFS distances.
// A feature cell's motion goals are its floor neighbours, facing it (planners.py:283-295);
// goals facing floor or a counter are not valid (planners.py:117-128, counter_goals = []).
extern "C" int pth_overcooked_layout_init(pth_overcooked_layout* L) {
  PTH_CHECK_ARG(L != nullptr, "NULL layout");
  const int W = L->width, Hh = L->height;
  PTH_CHECK_ARG(W >= 3 && Hh >= 3 && W * Hh <= PTH_OC_MAX_CELLS, "grid must be 3x3 .. 128 cells");
  PTH_CHECK_ARG(L->num_items >= 1 && L->num_items <= 255 && L->cook_time >= 0 && L->cook_time <= 250,
                "bad num_items / cook_time");
  PTH_CHECK_ARG(L->horizon >= 1 && L->horizon <= 65535, "horizon must fit 16 bits");
  PTH_CHECK_ARG(L->ego_agent_idx == 0 || L->ego_agent_idx == 1, "ego_agent_idx must be 0 or 1");
  const int cells = W * Hh;
  static const int DX[4] = {0, 0, 1, -1}, DY[4] = {-1, 1, 0, 0}, OPP[4] = {1, 0, 3, 2};
  L->n_pots = L->n_counters = 0;
  for (int c = 0; c < PTH_OC_MAX_CELLS; ++c) {
    L->slot[c] = 255;
    L->wall[c] = 0;
  }
  for (int y = 0; y < Hh; ++y)
    for (int x = 0; x < W; ++x) {
      const int c = y * W + x, t = L->terrain[c];
      if (t > PTH_OC_SERVE) {
        pth_set_error("pth_overcooked_layout_init: unknown terrain code %d (tomato layouts are not supported)", t);
        return PTH_ENOSUP;
      }
      if ((x == 0 || y == 0 || x == W - 1 || y == Hh - 1) && t == PTH_OC_FLOOR) {
        pth_set_error("pth_overcooked_layout_init: the grid border must not be floor");
        return PTH_EINVAL;
      }
      if (t == PTH_OC_COUNTER) {
        if (L->n_counters >= PTH_OC_MAX_COUNTERS) {
          pth_set_error("pth_overcooked_layout_init: more than %d counters", PTH_OC_MAX_COUNTERS);
          return PTH_ENOSUP;
        }
        L->slot[c] = (uint8_t)L->n_counters++;
      } else if (t == PTH_OC_POT) {
        if (L->n_pots >= PTH_OC_MAX_POTS) {
          pth_set_error("pth_overcooked_layout_init: more than %d pots", PTH_OC_MAX_POTS);
          return PTH_ENOSUP;
        }
        L->pot_x[L->n_pots] = (uint8_t)x;
        L->pot_y[L->n_pots] = (uint8_t)y;
        L->slot[c] = (uint8_t)L->n_pots++;
      }
    }
  for (int i = 0; i < 2; ++i) {
    const int x = L->start_x[i], y = L->start_y[i];
    PTH_CHECK_ARG(x > 0 && y > 0 && x < W - 1 && y < Hh - 1 && L->terrain[y * W + x] == PTH_OC_FLOOR,
                  "start positions must be floor cells");
  }
  PTH_CHECK_ARG(L->start_x[0] != L->start_x[1] || L->start_y[0] != L->start_y[1], "players start on the same cell");
  for (int y = 1; y < Hh - 1; ++y)
    for (int x = 1; x < W - 1; ++x)
      for (int d = 0; d < 4; ++d)
        if (L->terrain[(y + DY[d]) * W + x + DX[d]] != PTH_OC_FLOOR) L->wall[y * W + x] |= (uint8_t)(1 << d);
  memset(L->static_delta, 0, sizeof(L->static_delta));
  memset(L->pot_dist, 255, sizeof(L->pot_dist));
  // BFS from every node
  static thread_local int dist[PTH_OC_MAX_CELLS * 4], queue[PTH_OC_MAX_CELLS * 4];
  const int nodes = cells * 4;
  for (int s0 = 0; s0 < nodes; ++s0) {
    if (L->terrain[s0 >> 2] != PTH_OC_FLOOR) continue;
    for (int i = 0; i < nodes; ++i) dist[i] = -1;
    int qh = 0, qt = 0;
    dist[s0] = 0;
    queue[qt++] = s0;
    while (qh < qt) {
      const int u = queue[qh++];
      const int c = u >> 2, o = u & 3, x = c % W, y = c / W;
      for (int a = 0; a < 4; ++a) {  // stay / interact are self loops
        int nx = x + DX[a], ny = y + DY[a];
        if (L->terrain[ny * W + nx] != PTH_OC_FLOOR) {
          nx = x;
          ny = y;
        }
        const int v = (ny * W + nx) * 4 + a;
        if (dist[v] < 0) {
          dist[v] = dist[u] + 1;
          queue[qt++] = v;
        }
      }
      (void)o;
    }
    // closest feature of each static class, features in scan order, goals in N,S,E,W order,
    // strictly-smaller wins (planners.py:250-272)
    const int sx = (s0 >> 2) % W, sy = (s0 >> 2) / W;
    const int klass[3] = {PTH_OC_ONION, PTH_OC_DISH, PTH_OC_SERVE};
    int best_static[3] = {-1, -1, -1}, bd_static[3] = {1 << 30, 1 << 30, 1 << 30};
    for (int c = 0; c < cells; ++c) {
      const int t = L->terrain[c];
      if (t == PTH_OC_FLOOR || t == PTH_OC_COUNTER) continue;
      int mind = 1 << 30;
      for (int d = 0; d < 4; ++d) {
        const int ax = c % W + DX[d], ay = c / W + DY[d];
        if (ax < 0 || ay < 0 || ax >= W || ay >= Hh || L->terrain[ay * W + ax] != PTH_OC_FLOOR) continue;
        const int g = dist[(ay * W + ax) * 4 + OPP[d]];
        if (g >= 0 && g < mind) mind = g;
      }
      if (mind == (1 << 30)) continue;
      if (t == PTH_OC_POT) {
        L->pot_dist[s0][L->slot[c]] = (uint8_t)(mind > 254 ? 254 : mind);
      } else {
        for (int k = 0; k < 3; ++k)
          if (t == klass[k] && mind < bd_static[k]) {
            bd_static[k] = mind;
            best_static[k] = c;
          }
      }
    }
    for (int k = 0; k < 3; ++k)
      if (best_static[k] >= 0) {
        L->static_delta[s0][k][0] = (int8_t)(best_static[k] % W - sx);
        L->static_delta[s0][k][1] = (int8_t)(best_static[k] / W - sy);
      }
  }
  return PTH_OK;
}

extern "C" int pth_env_overcooked_reset(pth_ctx* ctx, const pth_overcooked_layout* d_layout,
                                        pth_overcooked_state* d_state, float* d_obs, int64_t N,
                                        void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_layout && d_state, "NULL layout / state pointer");
  PTH_CHECK_ARG(((uintptr_t)d_layout % 8) == 0 && ((uintptr_t)d_state % 8) == 0 &&
                    (d_obs == nullptr || ((uintptr_t)d_obs % 16) == 0),
                "layout/state must be 8-byte aligned, obs 16-byte aligned");
  PTH_CHECK_ARG(N >= 0, "negative size");
  if (N == 0) return PTH_OK;
  overcooked_reset_kernel<<<pth_ceil_div(N, OC_TB), OC_TB, 0, (cudaStream_t)stream>>>(d_layout, d_state, d_obs, N);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

extern "C" int pth_env_overcooked_step(pth_ctx* ctx, const pth_overcooked_layout* d_layout,
                                       pth_overcooked_state* d_state, const uint8_t* d_ego_action,
                                       const uint8_t* d_alt_action, float* d_obs, float* d_reward,
                                       uint8_t* d_done, int64_t N, void* stream) {
  PTH_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  PTH_CHECK_ARG(d_layout && d_state && d_ego_action && d_alt_action && d_reward && d_done, "NULL device pointer");
  PTH_CHECK_ARG(((uintptr_t)d_layout % 8) == 0 && ((uintptr_t)d_state % 8) == 0 &&
                    (d_obs == nullptr || ((uintptr_t)d_obs % 16) == 0),
                "layout/state must be 8-byte aligned, obs 16-byte aligned");
  PTH_CHECK_ARG(N >= 0, "negative size");
  if (N == 0) return PTH_OK;
  overcooked_step_kernel<<<pth_ceil_div(N, OC_TB), OC_TB, 0, (cudaStream_t)stream>>>(
      d_layout, d_state, d_ego_action, d_alt_action, d_obs, d_reward, d_done, N);
  PTH_LAUNCH_CHECK();
  return PTH_OK;
}

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
