/*
 * pantheon_b200.h — C ABI of libpantheon_b200.so
 *
 * The drop-in boundary for the PPO rollout -> GAE -> update hot path of
 * Stanford-ILIAD/PantheonRL, rebuilt as hand-written sm_100a CUDA.
 *
 * The reference has no FFI: its "operator API" for this path is a duck-typed
 * Python interface (pantheonrl/common/agents.py:24-51 Agent.get_action/update,
 * agents.py:111-203 OnPolicyAgent) that reaches into Stable-Baselines3 objects
 * (policy.forward, rollout_buffer.add / compute_returns_and_advantage,
 * model.train).  Each entry point below cites the reference call site it
 * replaces.  The Python facade in pantheonrl_b200/ binds these with ctypes
 * (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 on success or a negative PTH_E* code;
 *     pth_last_error() gives a thread-local message for the last failure.
 *   - all pointers named d_* (or inside *_args structs unless stated) are
 *     DEVICE pointers owned by the caller (PyTorch tensors: tensor.data_ptr()).
 *     The library never allocates, frees or retains device memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     Calls are asynchronous with respect to the host unless stated.
 *   - layouts are time-major like SB3's RolloutBuffer: [T][N] with the env
 *     index contiguous.
 *   - all floating point is IEEE fp32, round-to-nearest, no fast-math; the
 *     evaluation order of every reduction is fixed (see DESIGN.md "numeric
 *     contract") so results are run-to-run and CPU-oracle bit-reproducible.
 */
#ifndef PANTHEON_B200_H
#define PANTHEON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTH_VERSION 100 /* 0.1.0 */

/* error codes */
#define PTH_OK 0
#define PTH_EINVAL (-1)   /* bad argument (null pointer, bad size, bad space) */
#define PTH_ECUDA (-2)    /* CUDA runtime error (message in pth_last_error) */
#define PTH_ENOSUP (-3)   /* configuration not supported by this build */
#define PTH_ENODEV (-4)   /* no usable sm_100 device */

typedef struct pth_ctx pth_ctx;

int pth_version(void);
const char* pth_last_error(void);

/* One ctx per (process, device).  Queries SM count / limits once.
 * Fails with PTH_ENODEV when no CUDA device is present: there is NO CPU
 * fallback anywhere in this library. */
int pth_ctx_create(int device, pth_ctx** out);
int pth_ctx_destroy(pth_ctx* ctx);
int pth_ctx_sm_count(const pth_ctx* ctx);
/* debug only: cudaStreamSynchronize(stream) and surface async errors */
int pth_sync_debug(pth_ctx* ctx, void* stream);

/* ------------------------------------------------------------------ */
/* Spaces and policy parameter layout                                  */
/* ------------------------------------------------------------------ */

#define PTH_MAX_OBS_SLOTS 96
/* bytes of a one-hot observation row in rollout buffers / update inputs: one byte per slot, padded to
 * 32, or to 96 for spaces with more than 32 slots (frame-stacked observations, wrappers.py:233-349) */
#define PTH_OBS_ROW_BYTES(obs_len) ((obs_len) <= 32 ? 32 : 96)
#define PTH_MAX_HEADS 4
#define PTH_HIDDEN 64 /* SB3 MlpPolicy default: pi=[64,64], vf=[64,64], tanh */

enum { PTH_OBS_ONEHOT = 0, PTH_OBS_BOX = 1 };

/* Observation / action description of one agent's policy.
 *   PTH_OBS_ONEHOT: obs is obs_len small integers (Discrete(n): obs_len=1,
 *     MultiDiscrete(nvec): obs_len=len(nvec)); features = concat of one-hots,
 *     width = sum(obs_nvec)   (SB3 preprocess_obs)
 *   PTH_OBS_BOX: obs is obs_len fp32 values, features = obs.
 * Action: n_heads independent categoricals (Discrete: 1 head;
 * MultiDiscrete([7,12]): 2 heads), logits concatenated (SB3
 * MultiCategoricalDistribution; Appendix A3). */
typedef struct pth_space {
  int32_t obs_kind;
  int32_t obs_len;
  int32_t obs_nvec[PTH_MAX_OBS_SLOTS];
  int32_t n_heads;
  int32_t head_n[PTH_MAX_HEADS];
} pth_space;

/* Flat fp32 parameter vector, SB3 registration order (Appendix A2):
 *   pi0.w[F][64] pi0.b[64] pi1.w[64][64] pi1.b[64]
 *   vf0.w[F][64] vf0.b[64] vf1.w[64][64] vf1.b[64]
 *   act.w[L][64] act.b[L]  val.w[1][64]  val.b[1]
 * F = feature width, L = sum(head_n).  The two first-layer matrices are stored
 * INPUT-MAJOR [F][64] — the transpose of torch's nn.Linear.weight — so that a
 * one-hot row gather reads 256 contiguous bytes; every other tensor keeps
 * torch's weight[out][in] layout.  (The Python facade transposes at the
 * state_dict boundary.) */
int pth_space_feature_dim(const pth_space* sp);
int pth_space_logit_dim(const pth_space* sp);
int64_t pth_policy_param_count(const pth_space* sp);
/* AdapPolicy (pantheonrl/algos/adap/policies.py:21-131): the same network whose two first layers take
 * `context_size` (<= 8) more inputs behind the features — `features = cat(features, context)` — so
 * the two first-layer matrices have F + context_size rows (context rows last, same input-major
 * layout) and every other tensor is unchanged. */
int64_t pth_adap_param_count(const pth_space* sp, int32_t context_size);

/* ------------------------------------------------------------------ */
/* a4: GAE / returns                                                   */
/* replaces RolloutBuffer.compute_returns_and_advantage called at      */
/* pantheonrl/common/agents.py:127-130 (partner) and by SB3            */
/* collect_rollouts for the ego (restated adap_learn.py:468-469)       */
/* ------------------------------------------------------------------ */

/* Dense buffer: rewards, values, episode_starts, advantages, returns are
 * [T][N] fp32; last_values, dones are [N] fp32.
 *   nnt_t   = 1 - episode_starts[t+1]        (t < T-1),  1 - dones  (t = T-1)
 *   delta_t = rewards[t] + gamma*V_{t+1}*nnt_t - values[t]
 *   A_t     = delta_t + gamma*lambda*nnt_t*A_{t+1};   R_t = A_t + values[t]
 * One thread per env walks t = T-1..0 with a software-pipelined window of
 * loads; bit-exact with the sequential CPU recurrence.  20 B per (t, env).
 * variant: 0 = auto, 1 = register-window LDG kernel, 2 = TMA/mbarrier
 * smem-staged kernel, 3 = time-parallel warp affine scan (small N; agrees with
 * the sequential recurrence to rounding, not bit-exact). */
int pth_gae_f32(pth_ctx* ctx, const float* d_rewards, const float* d_values,
                const float* d_episode_starts, const float* d_last_values,
                const float* d_dones, float* d_advantages, float* d_returns,
                int64_t T, int64_t N, double gamma, double gae_lambda,
                int variant, void* stream);

/* Ragged buffer (vectorised partner, DESIGN.md "partner buffer"): same arrays
 * with leading dimension Tcap, per-env valid prefix length d_count[n] <= Tcap.
 * Bootstrap follows agents.py:127-129 exactly: last_values = the value stored
 * with the env's LAST recorded decision, dones = d_last_done[n] (the flag
 * latched by the last update(), agents.py:197).  Rows >= count are untouched. */
int pth_gae_ragged_f32(pth_ctx* ctx, const float* d_rewards,
                       const float* d_values, const float* d_episode_starts,
                       const int32_t* d_count, const float* d_last_done,
                       float* d_advantages, float* d_returns, int64_t Tcap,
                       int64_t N, double gamma, double gae_lambda,
                       void* stream);

/* ------------------------------------------------------------------ */
/* a8/a9: game rules as standalone step kernels (one thread per env)   */
/* ------------------------------------------------------------------ */

/* RPSEnv.multi_step, pantheonrl/envs/rpsgym/rps.py:41-45.
 * actions int32 [N] in {0,1,2}; rewards fp32 [N] (ego, alt); done always. */
int pth_env_rps_step(pth_ctx* ctx, const int32_t* d_ego_action,
                     const int32_t* d_alt_action, float* d_ego_reward,
                     float* d_alt_reward, int64_t N, void* stream);

/* Liar's Dice state, one 32-byte record per env (liar.py:45-102):
 *   hands[0..5] = ego histogram, hands[6..11] = partner histogram
 *   hist[i] = face | (count << 3), newest bid first (liar.py:82), hist_len bids
 */
#define PTH_LIAR_SIDES 6
#define PTH_LIAR_DICE 6
#define PTH_LIAR_MAX_MOVES 12
#define PTH_LIAR_OBS_LEN 30
typedef struct pth_liar_state {
  uint8_t hands[12];
  uint8_t hist[12];
  uint8_t hist_len;
  uint8_t pad[7];
} pth_liar_state;

/* LiarEnv.multi_reset (liar.py:97-102) with the dice supplied by the caller's
 * counter-based RNG: hands drawn from Philox4x32-10(seed, stream ENV) at
 * (env0 + n, tick); also draws ego_first = u < probegostart
 * (multiagentenv.py:325).  obs out: u8 [N][32] observation of the mover. */
int pth_env_liar_reset(pth_ctx* ctx, pth_liar_state* d_state,
                       uint8_t* d_ego_first, uint8_t* d_obs, int64_t N,
                       uint64_t seed, uint32_t tick, int64_t env0,
                       float probegostart, void* stream);

/* LiarEnv.player_step (liar.py:77-83) for the player given per env by
 * d_is_ego[n] (1 = ego_step, 0 = alt_step): sanitize_action (liar.py:58-67),
 * eval_bluff (liar.py:69-75).  action u8 [N][2] (face 0..6, count 0..11).
 * Outputs: obs of the OTHER player u8 [N][32], rewards (ego, alt) fp32, done. */
int pth_env_liar_step(pth_ctx* ctx, pth_liar_state* d_state,
                      const uint8_t* d_is_ego, const uint8_t* d_action,
                      uint8_t* d_obs, float* d_ego_reward, float* d_alt_reward,
                      uint8_t* d_done, int64_t N, void* stream);

/* ------------------------------------------------------------------ */
/* a10: Overcooked gridworld behind OvercookedMultiEnv                 */
/* replaces overcookedgym/overcooked.py:51-98 (multi_step/multi_reset) */
/* -> overcooked_ai_py/mdp/overcooked_env.py:77-121 (step/reset/done)  */
/* -> mdp/overcooked_mdp.py:643-846 (get_state_transition) and         */
/* :1077-1170 (featurize_state), planning/planners.py:250-295          */
/* (MotionPlanner.min_cost_to_feature with NO_COUNTERS_PARAMS)         */
/* ------------------------------------------------------------------ */

#define PTH_OC_MAX_CELLS 128    /* width * height (largest reference layout: corridor 14 x 9) */
#define PTH_OC_MAX_POTS 4
#define PTH_OC_MAX_COUNTERS 64
#define PTH_OC_OBS 62           /* featurize_state width (29 + 29 + 2 + 2) */
#define PTH_OC_ROW 64           /* floats per observation row in the rollout buffers */
enum { PTH_OC_FLOOR = 0, PTH_OC_COUNTER = 1, PTH_OC_ONION = 2, PTH_OC_POT = 3, PTH_OC_DISH = 4, PTH_OC_SERVE = 5 };

/* One layout + the env constants of OvercookedMultiEnv, plus lookup tables derived from
 * them.  Fill the first block, call pth_overcooked_layout_init (HOST, no device work),
 * then copy the struct to device memory and pass that pointer as d_layout.
 * Onion-only layouts (every layout trainer.py accepts except mdp_test / simple_tomato,
 * on which the reference's featurize_state itself raises once a tomato is held). */
typedef struct pth_overcooked_layout {
  /* ---- inputs */
  int32_t width, height;
  int32_t cook_time, num_items, delivery_reward, horizon;   /* layout file + DEFAULT_ENV_PARAMS (overcooked.py:18-20) */
  int32_t rew_placement_in_pot, rew_dish_pickup, rew_soup_pickup; /* rew_shaping_params (overcooked.py:21-28) */
  int32_t ego_agent_idx;                                    /* OvercookedMultiEnv(ego_agent_idx=) */
  int32_t start_x[2], start_y[2];                           /* start_player_positions */
  uint8_t terrain[PTH_OC_MAX_CELLS];                        /* row-major y*width+x, PTH_OC_* codes */
  /* ---- derived by pth_overcooked_layout_init */
  int32_t n_pots, n_counters;
  uint8_t slot[PTH_OC_MAX_CELLS];        /* counter index of an X cell, pot index of a P cell, 255 otherwise */
  uint8_t wall[PTH_OC_MAX_CELLS];        /* bit d: the neighbour in direction d (N,S,E,W) is not floor */
  uint8_t pot_x[PTH_OC_MAX_POTS], pot_y[PTH_OC_MAX_POTS];
  /* per MotionPlanner node = cell*4 + orientation: */
  int8_t static_delta[PTH_OC_MAX_CELLS * 4][3][2];  /* (dx, dy) to the closest onion dispenser / dish dispenser / serving cell; (0,0) if none */
  uint8_t pot_dist[PTH_OC_MAX_CELLS * 4][PTH_OC_MAX_POTS]; /* plan length to pot p's closest valid motion goal; 255 = unreachable */
} pth_overcooked_layout;

/* HOST: validates the grid and fills the derived tables (BFS over the (position,
 * orientation) motion graph).  PTH_ENOSUP for tomato layouts / too many cells, pots or counters. */
int pth_overcooked_layout_init(pth_overcooked_layout* host_layout);

/* Per-env game state, 40 bytes.  Objects on counters: 2 bits per counter slot
 * (bit in ctr_lo | bit in ctr_hi << 1): 0 none, 1 onion, 2 soup (always a finished
 * onion soup), 3 dish.  held: same codes.  Pot p: pot_n onions, pot_t cook time. */
typedef struct pth_overcooked_state {
  uint64_t ctr_lo, ctr_hi;
  uint8_t px[2], py[2], po[2], held[2];
  uint8_t pot_n[PTH_OC_MAX_POTS], pot_t[PTH_OC_MAX_POTS];
  uint16_t t;
  uint8_t pad[6];
} pth_overcooked_state;

/* OvercookedMultiEnv.multi_reset: standard start state; obs out fp32 [N][2][PTH_OC_ROW]
 * (ego's row first, then the partner's; columns 62, 63 are zero). */
int pth_env_overcooked_reset(pth_ctx* ctx, const pth_overcooked_layout* d_layout,
                             pth_overcooked_state* d_state, float* d_obs, int64_t N, void* stream);
/* OvercookedMultiEnv.multi_step for N envs: actions u8 [N] in Action.INDEX_TO_ACTION
 * order (N, S, E, W, stay, interact) for the ego and the partner; reward (same for both
 * agents: sparse + shaped) fp32 [N]; done u8 [N] (horizon).  Envs are NOT auto-reset.
 * d_obs as above for the next state. */
int pth_env_overcooked_step(pth_ctx* ctx, const pth_overcooked_layout* d_layout,
                            pth_overcooked_state* d_state, const uint8_t* d_ego_action,
                            const uint8_t* d_alt_action, float* d_obs, float* d_reward,
                            uint8_t* d_done, int64_t N, void* stream);

/* ------------------------------------------------------------------ */
/* a1 + a2: policy forward + sample + log-prob + value (+ buffer row)  */
/* replaces util.action_from_policy (util.py:63-81) ->                 */
/* ActorCriticPolicy.forward and RolloutBuffer.add (agents.py:162-179) */
/* ------------------------------------------------------------------ */

/* B independent observations.  obs: u8 [B][obs_stride] (ONEHOT) or fp32
 * [B][obs_stride] (BOX).  Sampling: inverse-CDF on one Philox uniform per
 * head, counter (idx0 + b, tick, slot), stream = rng_stream.  If
 * d_action_in != NULL no sampling happens and log-prob/entropy are evaluated
 * for the given actions (ActorCriticPolicy.evaluate_actions).
 * Outputs (any may be NULL): action u8 [B][4], value fp32 [B], logp fp32 [B],
 * entropy fp32 [B], logits fp32 [B][L]. */
typedef struct pth_forward_args {
  const pth_space* space;      /* HOST pointer */
  const float* d_params;
  const void* d_obs;
  int64_t obs_stride;          /* elements per row */
  int64_t B;
  uint64_t seed;
  uint32_t rng_stream;
  uint32_t tick;
  uint32_t slot;
  int64_t idx0;
  const uint8_t* d_action_in;
  uint8_t* d_action;
  float* d_value;
  float* d_logp;
  float* d_entropy;
  float* d_logits;
  /* Reference-RNG compatibility (N = 1 facade): when sampling and d_race != NULL, head h picks
   * argmax_i (p_i / q_i) over its own logits, q = d_race[b][L] exponential(1) draws supplied by the
   * host — exactly what torch.multinomial(probs, 1) computes from its generator
   * (Categorical.sample, the reference's sampling path: util.py:63-81 -> SB3 forward), so a host that
   * draws q from torch's generator in the reference's order reproduces the reference's action stream
   * whenever the logits agree.  NULL: inverse-CDF sampling on the Philox stream (default). */
  const float* d_race;
  /* AdapPolicy.forward / evaluate_actions (adap/policies.py:86-131): context_size > 0 selects the
   * AdapPolicy parameter layout; d_context is fp32 [B][context_size] with row stride
   * context_stride floats (0: one context for the whole batch, `self.context.repeat(B, 1)`). */
  int32_t context_size;
  const float* d_context;
  int64_t context_stride;
  /* ModularPolicy.forward / evaluate_actions (modular/policies.py:273-290, 364-383): num_partners > 0
   * selects the ModularPolicy parameter layout (pth_modular_param_count) and partner module
   * partner_idx composes the outputs: logits = main + partner, value = main + partner. */
  int32_t num_partners;
  int32_t partner_idx;
  /* AdapPolicyMult: adap_mult != 0 with context_size > 0 (parameter layout of pth_adap_mult_param_count). */
  int32_t adap_mult;
} pth_forward_args;
int pth_policy_forward(pth_ctx* ctx, const pth_forward_args* args, void* stream);
/* debug / parity only: y[i] = f(x[i]) with the library's exp (which = 0), log (1, positive
 * normal inputs) or tanh (2) — the polynomial kernels of the numeric contract. */
int pth_debug_math(pth_ctx* ctx, int which, const float* d_x, float* d_y, int64_t n, void* stream);
/* measurement only: the FP32 FFMA rate the MLP kernels are bounded by (forward and
 * update are FP32-compute bound).  One launch of `ctas` x 1024 threads, each running 8 independent
 * register-resident fmaf chains for `iters` rounds: ctas * 1024 * iters * 8 FFMA = 2x that in FLOP.
 * The caller times it with CUDA events on `stream`; d_sink [ctas * 1024] floats keeps the chains live. */
int pth_debug_ffma_peak(pth_ctx* ctx, float* d_sink, int32_t ctas, int32_t iters, void* stream);

/* ------------------------------------------------------------------ */
/* a1+a2+a3+a7+a8/a9 fused: T-tick rollout with on-device envs         */
/* replaces the Python loop OnPolicyAlgorithm.collect_rollouts ->      */
/* MultiAgentEnv.step (multiagentenv.py:172-215) -> partner.get_action */
/* / update (agents.py:111-203) -> n_step / n_reset                    */
/* ------------------------------------------------------------------ */

enum { PTH_ENV_RPS = 0, PTH_ENV_LIAR = 1, PTH_ENV_OVERCOOKED = 2 };

/* rollout buffer of one learner; all arrays [Tcap][N].  Observation rows: 32 B
 * (PTH_OBS_ONEHOT: one byte per slot) or PTH_OC_ROW fp32 = 256 B (PTH_OBS_BOX).
 * Ego: Tcap = T, dense (count unused, may be NULL).
 * Partner: Tcap >= 2*T + 1 (turn-based) or T (simultaneous), ragged; count[n] = rows of this
 * rollout that are complete (trainable).  Turn-based games: a row still waiting for the ego's next
 * move when the rollout ends is not counted; it stays at row count[n], carry flag bit2 is set, and
 * the next pth_rollout_run on the SAME buffer starts by moving it to row 0 (so no reward is lost at
 * a rollout boundary; the reference's lazily training partner sees it too, agents.py:126, 198). */
typedef struct pth_buffer {
  uint8_t* d_obs;            /* [Tcap][N][32] u8, or [Tcap][N][64] fp32 for Box spaces */
  uint8_t* d_actions;        /* [Tcap][N][4]  */
  float* d_rewards;          /* [Tcap][N] */
  float* d_values;           /* [Tcap][N] */
  float* d_logp;             /* [Tcap][N] */
  float* d_episode_starts;   /* [Tcap][N] */
  int32_t* d_count;          /* [N] or NULL */
  int64_t Tcap;
} pth_buffer;

/* persistent per-env driver state carried across rollouts (the fields of
 * MultiAgentEnv, multiagentenv.py:58-66, plus agents.py:97 latches) */
typedef struct pth_env_carry {
  float* d_ego_last_start;     /* [N] ego _last_episode_starts            */
  float* d_alt_last_done;      /* [N] partner _last_episode_starts latch  */
  float* d_total_rew;          /* [2][N] total_rews                       */
  uint8_t* d_flags;            /* [N] bit0 ego_moved, bit1 should_update, bit2 open partner row carried */
  void* d_game_state;          /* PTH_ENV_LIAR: pth_liar_state[N]; PTH_ENV_OVERCOOKED: pth_overcooked_state[N]; RPS: NULL */
  float* d_ego_last_value;     /* [N] out: V(obs_T) bootstrap for the ego  */
  float* d_ego_last_done;      /* [N] out: done flag after the last tick   */
  float* d_ep_stats;           /* [4] += {episodes, sum ego ep reward, sum ep len, partner decisions} or NULL */
  float* d_alt_boot_done;      /* [N] out or NULL: the `dones` argument of the partner's compute_returns_and_advantage
                                * (agents.py:127-130) for this rollout's closed rows: the partner's done latch, or —
                                * turn-based games, when the partner's latest row is still open — the
                                * episode_start stored with that open row */
} pth_env_carry;

typedef struct pth_rollout_args {
  int32_t env_kind;
  int32_t partner_records;     /* 1: partner is a learner (OnPolicyAgent); 0: StaticPolicyAgent */
  const pth_space* space;      /* HOST pointer; ego and partner share spaces (getDummyEnv, multiagentenv.py:72-79) */
  const float* d_ego_params;
  const float* d_alt_params;   /* == d_ego_params for self-play */
  pth_buffer ego;
  pth_buffer alt;
  pth_env_carry carry;
  int64_t N;
  int64_t T;
  int64_t env0;                /* global index of env 0 of this shard (multi-GPU) */
  uint64_t seed;
  uint32_t tick0;              /* global tick of the first tick of this rollout */
  float probegostart;          /* TurnBasedEnv.probegostart */
  int32_t first_rollout;       /* 1: envs are reset before tick 0 (SB3 _setup_learn) */
  const pth_overcooked_layout* d_layout; /* PTH_ENV_OVERCOOKED: DEVICE copy of an initialised layout; else NULL */
} pth_rollout_args;
int pth_rollout_run(pth_ctx* ctx, const pth_rollout_args* args, void* stream);

/* ------------------------------------------------------------------ */
/* a5: PPO update                                                      */
/* replaces model.train() at agents.py:155 -> SB3 PPO.train            */
/* (restated pantheonrl/algos/adap/adap_learn.py:229-347)              */
/* ------------------------------------------------------------------ */

/* Keyed bijection on [0, M): perm[e][i] for epochs e in [0, n_epochs).
 * Stands in for np.random.permutation (SB3 RolloutBuffer.get): 4-round
 * Feistel on ceil-pow2 domain with cycle walking, keyed by Philox(seed,
 * stream SHUFFLE, epoch_counter0 + e).  d_perm int32 [n_epochs][M]. */
int pth_perm_feistel(pth_ctx* ctx, int32_t* d_perm, int64_t M, int32_t n_epochs,
                     uint64_t seed, uint32_t stream_id, uint32_t epoch0,
                     void* stream);

/* Compact the valid (row, env) cells of a ragged buffer into flat sample
 * offsets row*N + env, env-major order (SB3 swap_and_flatten): d_index int32
 * [>= sum(count)], *d_total = sum(count).  For a dense buffer pass
 * d_count = NULL and every env gets T rows. */
int pth_index_build(pth_ctx* ctx, const int32_t* d_count, int64_t T, int64_t N,
                    int32_t* d_index, int32_t* d_total, void* d_workspace,
                    void* stream);
int64_t pth_index_workspace_bytes(int64_t N);

#define PTH_UPDATE_FLAG_WORDS 16384 /* per rank: 4 uint32 per minibatch of a sharded launch ({mean, tag}, {std, tag}) */

typedef struct pth_update_args {
  const pth_space* space;   /* HOST pointer */
  float* d_params;          /* [P] in/out */
  float* d_adam_m;          /* [P] in/out */
  float* d_adam_v;          /* [P] in/out */
  int64_t adam_step;        /* optimiser steps already taken */
  /* flat sample arrays indexed by offset = d_index[perm[...]] */
  const uint8_t* d_obs;     /* [*][32] u8 (ONEHOT) ; BOX: fp32 rows via d_obs_f32 */
  const float* d_obs_f32;   /* BOX: [*][obs_stride] fp32 rows (16-byte aligned), else NULL */
  int64_t obs_stride;       /* BOX: floats per row (multiple of 4, >= obs_len) */
  const uint8_t* d_actions; /* [*][4] */
  const float* d_old_logp;
  const float* d_advantages;
  const float* d_returns;
  int64_t rec_stride;       /* 0: separate arrays (32 B obs rows, 4 B others); else common byte stride (packed records) */
  const int32_t* d_index;   /* [M] sample -> flat offset, or NULL = identity */
  const int32_t* d_perm;    /* [n_epochs][M] */
  int64_t M;                /* samples in the buffer */
  int64_t batch_size;       /* SB3 batch_size; last minibatch may be short */
  int32_t n_epochs;
  float learning_rate, clip_range, ent_coef, vf_coef, max_grad_norm;
  float adam_beta1, adam_beta2, adam_eps;
  int32_t normalize_advantage;
  int32_t grid_ctas;        /* 0 = auto (pth_update_grid); tests pin it to compare with the oracle */
  /* Multi-GPU sharded update (world > 1): every rank holds the SAME sample arrays
   * (the all-gathered ego stream) and the same perm; tile t of a minibatch is
   * computed by rank t mod world; after the local ordered reduction each rank
   * stores its gradient sums into every peer's exchange buffer over NVLink as 8-byte
   * {value, epoch tag} pairs ("LL" protocol: the reader polls the pair itself until the
   * tag is the minibatch's epoch — no fence, no flag, no grid-wide barrier; slice c of rank
   * r only ever meets slice c of the other ranks), and all ranks add the per-rank sums in
   * rank order inside the same persistent kernel — replicas stay bit-identical.
   * d_peer_xbuf[r]: rank r's exchange buffer (>= pth_update_xbuf_bytes, zero-initialised
   * once) mapped in THIS process (HOST array of `world` device pointers).  d_peer_flags[r]:
   * rank r's advantage-statistics exchange (PTH_UPDATE_FLAG_WORDS uint32, zero-initialised once):
   * the per-minibatch {mean, std} of the advantages are computed by ONE rank each (minibatch id mod
   * world) and stored into every rank's array as {value, launch tag} pairs, so a launch may hold at
   * most PTH_UPDATE_FLAG_WORDS / 4 minibatches (n_epochs * n_minibatches) when world > 1.
   * flag_epoch: value of the monotonic epoch counter before this launch (launches add
   * n_epochs * n_minibatches; it must be the same on every rank). */
  int32_t world, rank;
  void* const* peer_xbuf;
  void* const* peer_flags;
  uint32_t flag_epoch;
  void* d_workspace;        /* pth_update_workspace_bytes() */
  int64_t workspace_bytes;
  float* d_stats;           /* [n_epochs*n_minibatch][8]: pg_loss, value_loss, entropy_loss, approx_kl, clip_frac, loss, grad_norm, n */
  /* loss_kind PTH_LOSS_BC: behaviour cloning on (obs, action) pairs instead of PPO
   * (pantheonrl/algos/bc.py:270-315): loss = -mean(log_prob) - ent_coef * mean(entropy)
   * + l2_weight * sum(theta^2) / 2; old_logp / advantages / returns are not read (pass any
   * valid pointers), vf_coef should be 0 and max_grad_norm +inf (bc.py does not clip).
   * stats columns then hold: neglogp, value_loss, -entropy, prob_true_act, 0, loss
   * (without the l2 term), grad_norm, n. */
  int32_t loss_kind;
  float l2_weight;
  /* AdapPolicy / ADAP.train (pantheonrl/algos/adap/adap_learn.py:229-347, adap/util.py:97-131).
   * context_size > 0: the parameters are an AdapPolicy's (pth_adap_param_count) and d_context holds
   * the context stored with every sample, fp32 [*][context_size], indexed by the same flat offset as
   * the other sample arrays (AdapAgent.get_action stores obs ++ context: adap/agent.py:121-124).
   * Needs rec_stride == 0, world == 1, at most 32 observation slots; workspace from
   * pth_adap_workspace_bytes.
   * loss_kind PTH_LOSS_ADAP adds `context_loss_coeff * context_loss` to PPO's loss.  Per minibatch
   * id = epoch * n_minibatches + m (B samples) the caller supplies the random draws of
   * get_context_kl_loss: d_ctx_states[id][num_state_samples] = positions inside the minibatch of the
   * sampled states (`th.randperm(B)[:num_state_samples]`; only the first min(num_state_samples, B)
   * are read) and d_ctx_draws[id][num_context_samples][context_size] = the sampled contexts
   * (`SAMPLERS[context_sampler]`).  The policy tower is evaluated on every (state, context) pair,
   * context_loss = mean over the K (K - 1) / 2 context pairs of mean_states exp(-KL(dist_a || dist_b)),
   * and its gradient flows into the policy tower and the action head.  d_stats column 5 (loss)
   * includes the term; d_ctx_loss (optional, [n_epochs * n_minibatch]) receives the context loss
   * itself (`train/context_kl_loss`, adap_learn.py:358). */
  int32_t context_size;
  const float* d_context;
  float context_loss_coeff;
  int32_t num_context_samples;  /* K: 2 .. 16 */
  int32_t num_state_samples;    /* S */
  const int32_t* d_ctx_states;
  const float* d_ctx_draws;
  float* d_ctx_loss;
  /* ModularAlgorithm.train (pantheonrl/algos/modular/learn.py:221-351) on a ModularPolicy
   * (modular/policies.py:23-396, defaults): loss_kind PTH_LOSS_MODULAR, ONE LAUNCH PER PARTNER PHASE.
   * Parameters: the MlpPolicy vector followed by num_partners blocks
   *   [w_pi0 64x64, b_pi0, w_pi1 64x64, b_pi1, w_vf0 64x64, b_vf0, w_vf1 64x64, b_vf1, w_act Lx64, b_act, w_val 64, b_val]
   * (matrices [out][in]; pth_modular_param_count): partner p's policy branch and value branch both read
   * the main policy tower's latent; logits = main + partner[p], value = main + partner[p].
   * The sample arrays are partner_idx's rollout buffer.  Loss = PPO's clipped surrogate / value /
   * entropy terms on the composed outputs + marginal_reg_coef * mean_b sum_l |mean_p softmax(main)_l -
   * mean_p softmax(main + partner[p])_l| (softmax over ALL logits jointly, learn.py:298-318), whose
   * gradient reaches every partner's policy branch.  The value branches of the OTHER partners get no
   * gradient and are left untouched (Adam skips tensors without a gradient); partner_vf_step = optimiser
   * steps partner_idx's value branch has taken so far (its own bias correction), adam_step = steps of
   * everything else.  d_stats column 5 includes the regulariser; d_ctx_loss (optional) receives it.
   * Needs rec_stride == 0, world == 1, context_size == 0, workspace from pth_modular_workspace_bytes and
   * d_modular_scratch (>= pth_modular_scratch_bytes, 16-byte aligned): per-CTA activation tiles of every
   * module, kept for the backward pass. */
  int32_t num_partners;
  int32_t partner_idx;
  int64_t partner_vf_step;
  float marginal_reg_coef;
  void* d_modular_scratch;
  int64_t modular_scratch_bytes;
  /* AdapPolicyMult (pantheonrl/algos/adap/policies.py:134-283, `trainer.py ... ADAP_MULT`): adap_mult != 0 with
   * context_size > 0.  Each tower is  x = tanh(W0 f + b0);  s = tanh(Ws x + bs) (64 -> 64 C);
   * y_j = x_j + sum_c s[j C + c] ctx_c;  out = tanh(W1 y + b1)  on the features WITHOUT the context.
   * Parameters (pth_adap_mult_param_count), per tower: W0 [F][64] input-major, b0, Ws [64 C][64], bs [64 C],
   * W1 [64][64], b1; then the heads.  loss_kind PTH_LOSS_PPO or PTH_LOSS_ADAP as for AdapPolicy; workspace from
   * pth_adap_mult_workspace_bytes, d_modular_scratch >= pth_adap_mult_scratch_bytes (per-CTA activation tiles). */
  int32_t adap_mult;
} pth_update_args;
#define PTH_LOSS_PPO 0
#define PTH_LOSS_BC 1
#define PTH_LOSS_ADAP 2
#define PTH_LOSS_MODULAR 3
int64_t pth_update_workspace_bytes(const pth_ctx* ctx, const pth_space* sp,
                                   int64_t M, int64_t batch_size);
int64_t pth_adap_workspace_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t context_size,
                                 int64_t M, int64_t batch_size);
int64_t pth_adap_mult_param_count(const pth_space* sp, int32_t context_size);
int64_t pth_adap_mult_workspace_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t context_size,
                                      int64_t M, int64_t batch_size);
int64_t pth_adap_mult_scratch_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t context_size);
int64_t pth_modular_param_count(const pth_space* sp, int32_t num_partners);
int64_t pth_modular_workspace_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t num_partners,
                                    int64_t M, int64_t batch_size);
int64_t pth_modular_scratch_bytes(const pth_ctx* ctx, const pth_space* sp, int32_t num_partners);
/* The random draws of ADAP on the library's Philox streams (the reference takes them from torch's
 * global generator: adap/util.py:42-94 SAMPLERS, :106 th.randperm; a host that wants the reference's
 * own stream draws them itself and passes them to pth_ppo_update — both are plain arrays).
 * For id in [0, n): d_draws[id][num_context_samples][context_size] = contexts from sampler
 * 0 "l2" (uniform in [-1, 1)^C scaled to the unit sphere), 1 "unit_square", 2 "positive_square",
 * 3 "categorical" (one-hot), 4 "natural_numbers" (context_size 1); and, when d_states != NULL,
 * d_states[id][num_state_samples] = the first num_state_samples images of a keyed permutation of
 * [0, B), B = size of minibatch id % n_minibatches of a buffer of M samples (-1 beyond B).
 * Counter = index0 + id: a caller advances index0 by n per launch (and uses another stream_id
 * for the per-episode context of a rollout: adap_learn.py:452-455, adap/agent.py:146-150). */
#define PTH_ADAP_SAMPLER_L2 0
#define PTH_ADAP_SAMPLER_UNIT_SQUARE 1
#define PTH_ADAP_SAMPLER_POSITIVE_SQUARE 2
#define PTH_ADAP_SAMPLER_CATEGORICAL 3
#define PTH_ADAP_SAMPLER_NATURAL_NUMBERS 4
int pth_adap_draw(pth_ctx* ctx, int32_t* d_states, float* d_draws, int64_t n, int64_t n_minibatches,
                  int64_t M, int64_t batch_size, int32_t num_state_samples, int32_t num_context_samples,
                  int32_t context_size, int32_t sampler, uint64_t seed, uint32_t stream_id,
                  uint32_t index0, void* stream);
/* Number of CTAs the persistent cooperative update kernel will run with for
 * this problem (it is part of the reduction contract: tile t of a minibatch is
 * summed by CTA t mod grid, CTAs are added in ascending order). */
int pth_update_grid(const pth_ctx* ctx, const pth_space* sp, int64_t M,
                    int64_t batch_size);
int64_t pth_update_xbuf_bytes(const pth_space* sp, int32_t world);
int pth_ppo_update(pth_ctx* ctx, const pth_update_args* args, void* stream);
/* debug only: when set to a device array of 32 int64 (zeroed by the caller), later
 * pth_ppo_update launches add CTA 0's per-phase clock64 sums to it (phase list in
 * csrc/pth_update.cu, PTH_PROF); NULL switches it off again.  Process-wide. */
int pth_debug_update_profile(void* d_clock_sums);

/* ------------------------------------------------------------------ */
/* e: multi-GPU exchange staging                                       */
/* ------------------------------------------------------------------ */

/* Pack one rank's ego transitions [T][N] into a contiguous record stream for a single all-gather
 * per rollout.  Record = observation row (obs_bytes: 32 for one-hot rows,
 * 4 * PTH_OC_ROW = 256 for Box rows) | action 4 | old_logp 4 | advantage 4 | return 4. */
#define PTH_PACKED_BYTES 48       /* one-hot spaces */
#define PTH_PACKED_BYTES_BOX 272  /* Box spaces     */
int pth_pack_transitions(pth_ctx* ctx, const uint8_t* d_obs, int32_t obs_bytes,
                         const uint8_t* d_actions, const float* d_logp,
                         const float* d_advantages, const float* d_returns,
                         int64_t count, uint8_t* d_packed, void* stream);
/* Same, but stores each record straight into every peer's gather buffer
 * (peer pointers mapped by the caller via CUDA IPC / symmetric memory): pack + all-gather in
 * one kernel over NVLink, no intermediate staging. d_peer_bufs: DEVICE array of
 * world pointers; records land at rank*count. */
int pth_pack_allgather_p2p(pth_ctx* ctx, const uint8_t* d_obs, int32_t obs_bytes,
                           const uint8_t* d_actions, const float* d_logp,
                           const float* d_advantages, const float* d_returns,
                           int64_t count, uint8_t* const* d_peer_bufs,
                           int32_t world, int32_t rank, void* stream);
/* The NCCL leg of the exchange inside the library (SURVEY.md 8b minimum export set), for hosts that
 * have no collective library of their own.  NCCL is bound at run time (dlopen of libnccl.so.2; a
 * process that already loaded one — PyTorch's — shares it); PTH_ENOSUP if it is not there.
 *   pth_comm_unique_id: rank 0 creates the 128-byte ncclUniqueId and hands it to the other ranks
 *                       by whatever channel the host has (file, socket, MPI, torch.distributed);
 *   pth_comm_init     : every rank joins (collective; one communicator per pth_ctx);
 *   pth_allgather_transitions: ncclAllGather of `bytes_per_rank` packed bytes from every rank into
 *                       d_gathered [world * bytes_per_rank] (rank r's records at r * bytes_per_rank),
 *                       enqueued on `stream`.  This is the ONE collective of the path per rollout. */
int pth_comm_unique_id(void* out128);
int pth_comm_init(pth_ctx* ctx, const void* unique_id128, int32_t world, int32_t rank);
int pth_comm_destroy(pth_ctx* ctx);
int pth_allgather_transitions(pth_ctx* ctx, const uint8_t* d_packed, int64_t bytes_per_rank,
                              uint8_t* d_gathered, void* stream);
/* The update reads a packed stream directly: point d_obs (Box: d_obs_f32) / d_actions /
 * d_old_logp / d_advantages / d_returns at offsets 0 / B / B+4 / B+8 / B+12 of the
 * first record (B = obs_bytes) and set rec_stride = B + 16. */

#ifdef __cplusplus
}
#endif
#endif /* PANTHEON_B200_H */
