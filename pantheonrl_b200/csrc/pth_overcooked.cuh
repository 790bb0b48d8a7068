// pth_overcooked.cuh — register-resident Overcooked gridworld (onion layouts) for
// the rollout megakernel and the standalone env kernels.
//
// Reference (paths under /root/reference/overcookedgym/):
//   overcooked.py:51-98                          OvercookedMultiEnv.multi_step / multi_reset
//   human_aware_rl/overcooked_ai/overcooked_ai_py/mdp/overcooked_env.py:77-121   step / reset / is_done
//   .../mdp/overcooked_mdp.py:643-675            get_state_transition
//   .../mdp/overcooked_mdp.py:677-756            resolve_interacts
//   .../mdp/overcooked_mdp.py:781-846            resolve_movement, collisions, _move_if_direction
//   .../mdp/overcooked_mdp.py:811-822            step_environment_effects
//   .../mdp/overcooked_mdp.py:1077-1170          featurize_state
//   .../planning/planners.py:250-295             min_cost_to_feature (tables built on the host,
//                                                pth_overcooked_layout_init in pth_envs.cu)
//
// State encoding (pth_overcooked_state): objects on counters are two bit planes over
// the layout's counter slots, pots are (items, cook time) byte pairs, everything a
// thread needs lives in 10 registers.  The layout tables are read from shared memory.
#pragma once
#include "pth_common.cuh"

struct OcRegs {
  uint64_t ctr_lo, ctr_hi;
  uint32_t pos;   // px0 | py0 << 8 | px1 << 16 | py1 << 24
  uint32_t oh;    // po0 | po1 << 8 | held0 << 16 | held1 << 24
  uint32_t pot_n, pot_t;  // 4 x u8
  uint32_t t;
};

enum { OC_HELD_NONE = 0, OC_HELD_ONION = 1, OC_HELD_SOUP = 2, OC_HELD_DISH = 3 };

__device__ __forceinline__ int oc_px(const OcRegs& s, int i) { return (s.pos >> (16 * i)) & 0xff; }
__device__ __forceinline__ int oc_py(const OcRegs& s, int i) { return (s.pos >> (16 * i + 8)) & 0xff; }
__device__ __forceinline__ int oc_po(const OcRegs& s, int i) { return (s.oh >> (8 * i)) & 0xff; }
__device__ __forceinline__ int oc_held(const OcRegs& s, int i) { return (s.oh >> (16 + 8 * i)) & 0xff; }
__device__ __forceinline__ void oc_set_held(OcRegs& s, int i, int v) {
  s.oh = (s.oh & ~(0xffu << (16 + 8 * i))) | ((uint32_t)v << (16 + 8 * i));
}
__device__ __forceinline__ void oc_set_po(OcRegs& s, int i, int v) {
  s.oh = (s.oh & ~(0xffu << (8 * i))) | ((uint32_t)v << (8 * i));
}
__device__ __forceinline__ void oc_set_pos(OcRegs& s, int i, int x, int y) {
  s.pos = (s.pos & ~(0xffffu << (16 * i))) | (((uint32_t)x | ((uint32_t)y << 8)) << (16 * i));
}
__device__ __forceinline__ int oc_byte(uint32_t w, int i) { return (w >> (8 * i)) & 0xff; }
__device__ __forceinline__ uint32_t oc_with_byte(uint32_t w, int i, int v) {
  return (w & ~(0xffu << (8 * i))) | ((uint32_t)v << (8 * i));
}

__device__ __forceinline__ void oc_load(const pth_overcooked_state* p, OcRegs& s) {
  const uint2* q = reinterpret_cast<const uint2*>(p);
  const uint2 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4];
  s.ctr_lo = (uint64_t)a.x | ((uint64_t)a.y << 32);
  s.ctr_hi = (uint64_t)b.x | ((uint64_t)b.y << 32);
  // c.x = px0 px1 py0 py1, c.y = po0 po1 held0 held1 (struct order: px[2] py[2] po[2] held[2])
  s.pos = (c.x & 0xffu) | (((c.x >> 16) & 0xffu) << 8) | (((c.x >> 8) & 0xffu) << 16) | ((c.x >> 24) << 24);
  s.oh = c.y;
  s.pot_n = d.x;
  s.pot_t = d.y;
  s.t = e.x & 0xffffu;
}

__device__ __forceinline__ void oc_store(pth_overcooked_state* p, const OcRegs& s) {
  uint2* q = reinterpret_cast<uint2*>(p);
  q[0] = make_uint2((uint32_t)s.ctr_lo, (uint32_t)(s.ctr_lo >> 32));
  q[1] = make_uint2((uint32_t)s.ctr_hi, (uint32_t)(s.ctr_hi >> 32));
  const uint32_t px0 = s.pos & 0xffu, py0 = (s.pos >> 8) & 0xffu, px1 = (s.pos >> 16) & 0xffu, py1 = s.pos >> 24;
  q[2] = make_uint2(px0 | (px1 << 8) | (py0 << 16) | (py1 << 24), s.oh);
  q[3] = make_uint2(s.pot_n, s.pot_t);
  q[4] = make_uint2(s.t & 0xffffu, 0u);
}

// OvercookedEnv.reset -> get_standard_start_state: start positions, facing NORTH, nothing held
__device__ __forceinline__ void oc_reset(const pth_overcooked_layout& L, OcRegs& s) {
  s.ctr_lo = 0;
  s.ctr_hi = 0;
  s.pos = (uint32_t)L.start_x[0] | ((uint32_t)L.start_y[0] << 8) | ((uint32_t)L.start_x[1] << 16) |
          ((uint32_t)L.start_y[1] << 24);
  s.oh = 0;
  s.pot_n = 0;
  s.pot_t = 0;
  s.t = 0;
}

__device__ __forceinline__ int oc_dx(int d) { return d == 2 ? 1 : (d == 3 ? -1 : 0); }  // N S E W
__device__ __forceinline__ int oc_dy(int d) { return d == 0 ? -1 : (d == 1 ? 1 : 0); }

// One joint action in PLAYER order (a[i] in 0..5: N, S, E, W, stay, interact).
// Returns done; reward = sparse + shaped (OvercookedMultiEnv gives it to both agents).
__device__ __forceinline__ bool oc_step(const pth_overcooked_layout& L, OcRegs& s, int a0, int a1,
                                        float& reward) {
  int rew = 0;
  // ---- resolve_interacts: pots holding a soup are counted before anything changes
  int nearly_ready = 0;
#pragma unroll
  for (int p = 0; p < PTH_OC_MAX_POTS; ++p) nearly_ready += (p < L.n_pots && oc_byte(s.pot_n, p) > 0) ? 1 : 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int act = i == 0 ? a0 : a1;
    if (act != 5) continue;
    const int o = oc_po(s, i);
    const int cell = (oc_py(s, i) + oc_dy(o)) * L.width + (oc_px(s, i) + oc_dx(o));
    const int tt = L.terrain[cell];
    const int held = oc_held(s, i);
    const int slot = L.slot[cell];
    if (tt == PTH_OC_COUNTER) {
      const uint64_t bit = 1ull << slot;
      const int obj = (int)((s.ctr_lo >> slot) & 1ull) | ((int)((s.ctr_hi >> slot) & 1ull) << 1);
      if (held != OC_HELD_NONE && obj == 0) {
        if (held & 1) s.ctr_lo |= bit;
        if (held & 2) s.ctr_hi |= bit;
        oc_set_held(s, i, OC_HELD_NONE);
      } else if (held == OC_HELD_NONE && obj != 0) {
        s.ctr_lo &= ~bit;
        s.ctr_hi &= ~bit;
        oc_set_held(s, i, obj);
      }
    } else if (tt == PTH_OC_ONION) {
      if (held == OC_HELD_NONE) oc_set_held(s, i, OC_HELD_ONION);
    } else if (tt == PTH_OC_DISH) {
      if (held == OC_HELD_NONE) {
        const int dishes_already = (oc_held(s, 0) == OC_HELD_DISH ? 1 : 0) + (oc_held(s, 1) == OC_HELD_DISH ? 1 : 0);
        oc_set_held(s, i, OC_HELD_DISH);
        const bool none_on_counters = (s.ctr_lo & s.ctr_hi) == 0ull;
        if (nearly_ready > dishes_already && none_on_counters) rew += L.rew_dish_pickup;
      }
    } else if (tt == PTH_OC_POT) {
      const int n = oc_byte(s.pot_n, slot), ct = oc_byte(s.pot_t, slot);
      if (held == OC_HELD_DISH) {
        if (n == L.num_items && ct >= L.cook_time) {
          oc_set_held(s, i, OC_HELD_SOUP);
          s.pot_n = oc_with_byte(s.pot_n, slot, 0);
          s.pot_t = oc_with_byte(s.pot_t, slot, 0);
          rew += L.rew_soup_pickup;
        }
      } else if (held == OC_HELD_ONION) {
        if (n < L.num_items) {  // empty pot, or a partly filled onion pot
          oc_set_held(s, i, OC_HELD_NONE);
          s.pot_n = oc_with_byte(s.pot_n, slot, n + 1);
          s.pot_t = oc_with_byte(s.pot_t, slot, 0);
          rew += L.rew_placement_in_pot;
        }
      }
    } else if (tt == PTH_OC_SERVE) {
      if (held == OC_HELD_SOUP) {
        oc_set_held(s, i, OC_HELD_NONE);
        rew += L.delivery_reward;
      }
    }
  }
  // ---- resolve_movement
  int nx[2], ny[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int act = i == 0 ? a0 : a1;
    nx[i] = oc_px(s, i);
    ny[i] = oc_py(s, i);
    if (act < 4) {
      oc_set_po(s, i, act);  // orientation changes even when the move is blocked or cancelled
      const int tx = nx[i] + oc_dx(act), ty = ny[i] + oc_dy(act);
      if (L.terrain[ty * L.width + tx] == PTH_OC_FLOOR) {
        nx[i] = tx;
        ny[i] = ty;
      }
    }
  }
  const bool same = nx[0] == nx[1] && ny[0] == ny[1];
  const bool swap = nx[0] == oc_px(s, 1) && ny[0] == oc_py(s, 1) && nx[1] == oc_px(s, 0) && ny[1] == oc_py(s, 0);
  if (!(same || swap)) {
    oc_set_pos(s, 0, nx[0], ny[0]);
    oc_set_pos(s, 1, nx[1], ny[1]);
  }
  // ---- step_environment_effects: full pots cook
#pragma unroll
  for (int p = 0; p < PTH_OC_MAX_POTS; ++p) {
    const int n = oc_byte(s.pot_n, p), ct = oc_byte(s.pot_t, p);
    if (p < L.n_pots && n == L.num_items && ct < L.cook_time) s.pot_t = oc_with_byte(s.pot_t, p, ct + 1);
  }
  s.t += 1;
  reward = (float)rew;
  return (int)s.t >= L.horizon;
}

// The 29-feature block of player i (featurize_state's per-player dictionary, in insertion order):
//  0-3 orientation one-hot | 4-6 held onion/soup/dish | 7,8 closest onion | 9,10 empty pot |
//  11-14 one_onion / two_onion pot (never filled by the reference: always 0) | 15,16 cooking pot |
//  17,18 ready pot | 19,20 closest dish | 21,22 closest soup (never reachable: always 0) |
//  23,24 serving | 25-28 wall N,S,E,W
template <typename Emit>
__device__ __forceinline__ void oc_player_block(const pth_overcooked_layout& L, const OcRegs& s, int i,
                                                Emit emit) {
  const int x = oc_px(s, i), y = oc_py(s, i), o = oc_po(s, i), held = oc_held(s, i);
  const int cell = y * L.width + x;
  const int node = cell * 4 + o;
#pragma unroll
  for (int d = 0; d < 4; ++d) emit(d, o == d ? 1.f : 0.f);
  emit(4, held == OC_HELD_ONION ? 1.f : 0.f);
  emit(5, held == OC_HELD_SOUP ? 1.f : 0.f);
  emit(6, held == OC_HELD_DISH ? 1.f : 0.f);
  const bool has_onion = held == OC_HELD_ONION, has_dish = held == OC_HELD_DISH;
  emit(7, has_onion ? 0.f : (float)L.static_delta[node][0][0]);
  emit(8, has_onion ? 0.f : (float)L.static_delta[node][0][1]);
  // pots by state: empty, cooking, ready — first pot (scan order) with the strictly smallest plan length
  int best[3] = {-1, -1, -1}, bd[3] = {255, 255, 255};
#pragma unroll
  for (int p = 0; p < PTH_OC_MAX_POTS; ++p) {
    if (p < L.n_pots) {
      const int n = oc_byte(s.pot_n, p), ct = oc_byte(s.pot_t, p);
      const int cls = n == 0 ? 0 : (n == L.num_items ? (ct < L.cook_time ? 1 : 2) : -1);
      const int dist = L.pot_dist[node][p];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (cls == c && dist < bd[c]) {
          bd[c] = dist;
          best[c] = p;
        }
    }
  }
  const float e0 = best[0] >= 0 ? (float)((int)L.pot_x[best[0] < 0 ? 0 : best[0]] - x) : 0.f;
  const float e1 = best[0] >= 0 ? (float)((int)L.pot_y[best[0] < 0 ? 0 : best[0]] - y) : 0.f;
  emit(9, e0);
  emit(10, e1);
  emit(11, 0.f); emit(12, 0.f); emit(13, 0.f); emit(14, 0.f);
  emit(15, best[1] >= 0 ? (float)((int)L.pot_x[best[1] < 0 ? 0 : best[1]] - x) : 0.f);
  emit(16, best[1] >= 0 ? (float)((int)L.pot_y[best[1] < 0 ? 0 : best[1]] - y) : 0.f);
  emit(17, best[2] >= 0 ? (float)((int)L.pot_x[best[2] < 0 ? 0 : best[2]] - x) : 0.f);
  emit(18, best[2] >= 0 ? (float)((int)L.pot_y[best[2] < 0 ? 0 : best[2]] - y) : 0.f);
  emit(19, has_dish ? 0.f : (float)L.static_delta[node][1][0]);
  emit(20, has_dish ? 0.f : (float)L.static_delta[node][1][1]);
  emit(21, 0.f); emit(22, 0.f);
  emit(23, (float)L.static_delta[node][2][0]);
  emit(24, (float)L.static_delta[node][2][1]);
  const int wl = L.wall[cell];
#pragma unroll
  for (int d = 0; d < 4; ++d) emit(25 + d, (wl >> d) & 1 ? 1.f : 0.f);
}

// Both agents' observations of one state, written feature-major into two shared-memory
// tiles (Xe: the ego's, Xa: the partner's; element (k, b) at k * lda + b).  Player block
// of player i lands at rows [0, 29) of its own observation and rows [29, 58) of the other's.
__device__ __forceinline__ void oc_write_obs(const pth_overcooked_layout& L, const OcRegs& s, float* Xe,
                                             float* Xa, int lda, int b) {
  const int e = L.ego_agent_idx;
  float* own[2] = {e == 0 ? Xe : Xa, e == 0 ? Xa : Xe};  // own[i]: observation of player i
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float* mine = own[i];
    float* other = own[1 - i];
    oc_player_block(L, s, i, [&](int k, float v) {
      mine[k * lda + b] = v;
      other[(29 + k) * lda + b] = v;
    });
  }
  const int x0 = oc_px(s, 0), y0 = oc_py(s, 0), x1 = oc_px(s, 1), y1 = oc_py(s, 1);
  own[0][58 * lda + b] = (float)(x1 - x0);
  own[0][59 * lda + b] = (float)(y1 - y0);
  own[1][58 * lda + b] = (float)(x0 - x1);
  own[1][59 * lda + b] = (float)(y0 - y1);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    own[i][60 * lda + b] = (float)x0;  // sic: both observations end with PLAYER 0's position
    own[i][61 * lda + b] = (float)y0;  // (overcooked_mdp.py:1150-1155)
    own[i][62 * lda + b] = 0.f;
    own[i][63 * lda + b] = 0.f;
  }
}
