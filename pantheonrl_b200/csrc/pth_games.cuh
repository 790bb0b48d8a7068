// pth_games.cuh — register-resident game rules for the two tabular envs.
//
// RPS:  pantheonrl/envs/rpsgym/rps.py:41-45 (RPSEnv.multi_step)
// Liar: pantheonrl/envs/liargym/liar.py:22-26 (randRoll), :53-56 (getObs),
//       :58-67 (sanitize_action), :69-75 (eval_bluff), :77-83 (player_step),
//       :97-102 (multi_reset); who starts: multiagentenv.py:325.
#pragma once
#include "pth_common.cuh"

// ------------------------------------------------------------------ RPS
__device__ __forceinline__ void pth_rps_outcome(int ego_a, int alt_a, float& r_ego,
                                                float& r_alt) {
  int o = (ego_a - alt_a + 3) % 3;
  o = (o == 2) ? -1 : o;
  r_ego = (float)o;
  r_alt = (float)(-o);
}

// ------------------------------------------------------------------ Liar's Dice
// Packed state: hands as twelve 4-bit histogram counts (ego in hand[0], partner
// in hand[1]); bids as bytes face | count << 3 in a 96-bit shift register with
// the newest bid in the low byte (the reference prepends, liar.py:82).
struct LiarRegs {
  uint32_t hand[2];
  uint32_t h0, h1, h2;
  uint32_t len;
};

#define PTH_LIAR_BLUFF_FACE 6
#define PTH_LIAR_BLUFF_COUNT 11

__device__ __forceinline__ uint32_t liar_hand_count(const LiarRegs& s, int player, int side) {
  return (s.hand[player] >> (4 * side)) & 0xfu;
}

__device__ __forceinline__ void liar_load(const pth_liar_state* p, LiarRegs& s) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  // a.x a.y a.z = hands[0..11], a.w b.x b.y = hist[0..11], b.z low byte = len
  uint32_t hw[3] = {a.x, a.y, a.z};
  s.hand[0] = 0;
  s.hand[1] = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    uint32_t v = (hw[i >> 2] >> (8 * (i & 3))) & 0xffu;
    s.hand[i / 6] |= (v & 0xfu) << (4 * (i % 6));
  }
  s.h0 = a.w;
  s.h1 = b.x;
  s.h2 = b.y;
  s.len = b.z & 0xffu;
}

__device__ __forceinline__ void liar_store(pth_liar_state* p, const LiarRegs& s) {
  uint32_t hw[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    uint32_t v = (s.hand[i / 6] >> (4 * (i % 6))) & 0xfu;
    hw[i >> 2] |= v << (8 * (i & 3));
  }
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(hw[0], hw[1], hw[2], s.h0);
  q[1] = make_uint4(s.h1, s.h2, s.len & 0xffu, 0u);
}

// Observation of `player` (0 ego, 1 partner) as 32 bytes: 6 histogram counts,
// then 12 (face, count) pairs newest first padded with DEFAULT = (6, 0)
// (liar.py:53-56), then 2 zero bytes.
__device__ __forceinline__ void liar_obs(const LiarRegs& s, int player, uint32_t (&w)[8]) {
  uint8_t b[32];
#pragma unroll
  for (int i = 0; i < 6; ++i) b[i] = (uint8_t)liar_hand_count(s, player, i);
  uint32_t hh[3] = {s.h0, s.h1, s.h2};
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    uint32_t bid = (hh[i >> 2] >> (8 * (i & 3))) & 0xffu;
    bool have = (uint32_t)i < s.len;
    b[6 + 2 * i] = have ? (uint8_t)(bid & 7u) : (uint8_t)6;
    b[7 + 2 * i] = have ? (uint8_t)(bid >> 3) : (uint8_t)0;
  }
  b[30] = 0;
  b[31] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) |
           ((uint32_t)b[4 * i + 3] << 24);
}

// player_step for `player`.  Returns done; rewards are (ego, partner).
__device__ __forceinline__ bool liar_step(LiarRegs& s, int player, int face, int count,
                                          float& r_ego, float& r_alt) {
  bool bluff = false;
  if (s.len != 0) {
    int top_count = (int)((s.h0 & 0xffu) >> 3);
    if (count <= top_count || face == PTH_LIAR_SIDES) bluff = true;
  } else if (face == PTH_LIAR_SIDES) {
    face = 0;
    count = 0;
  }
  if (bluff) {
    int side = (int)(s.h0 & 7u);
    int claimed = (int)((s.h0 & 0xffu) >> 3);
    int trueans = (int)liar_hand_count(s, 0, side) + (int)liar_hand_count(s, 1, side) - 1;
    bool was_bluff = claimed > trueans;
    bool caller_is_ego = (player == 0);
    bool ego_wins = (was_bluff == caller_is_ego);
    r_ego = ego_wins ? 1.0f : -1.0f;
    r_alt = ego_wins ? -1.0f : 1.0f;
    return true;
  }
  s.h2 = (s.h2 << 8) | (s.h1 >> 24);
  s.h1 = (s.h1 << 8) | (s.h0 >> 24);
  s.h0 = (s.h0 << 8) | ((uint32_t)face | ((uint32_t)count << 3));
  s.len += 1;
  r_ego = 0.f;
  r_alt = 0.f;
  return false;
}

// multi_reset + the TurnBasedEnv.n_reset coin.  Draw layout (stream ENV, index
// = global env, tick): flat draw d lives in Philox block slot_base + d / 4,
// lane d % 4; d = 0 is the who-starts uniform, d = 1..6 the ego dice, d = 7..12
// the partner dice; die = floor(u32 * 6 / 2^32).
__device__ __forceinline__ bool liar_reset(LiarRegs& s, uint64_t seed, uint64_t env,
                                           uint32_t tick, uint32_t slot_base,
                                           float probegostart) {
  uint32_t d[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    pth_u4 r = pth_philox(seed, PTH_STREAM_ENV, env, tick, slot_base + k);
    d[4 * k + 0] = r.x;
    d[4 * k + 1] = r.y;
    d[4 * k + 2] = r.z;
    d[4 * k + 3] = r.w;
  }
  bool ego_first = pth_u01(d[0]) < probegostart;
  s.hand[0] = 0;
  s.hand[1] = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    uint32_t side = pth_mulhi32(d[1 + i], 6u);
    s.hand[i / 6] += 1u << (4 * side);
  }
  s.h0 = s.h1 = s.h2 = 0;
  s.len = 0;
  return ego_first;
}
