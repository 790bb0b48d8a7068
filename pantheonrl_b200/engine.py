"""Vectorised on-device trainer: one ego PPO learner + one partner (PPO learner
or static / self-play policy) over N on-device env instances.

One ``iteration()`` = SB3's learn-loop body for BOTH agents
  collect_rollouts (pth_rollout_run)  ->  GAE (pth_gae_f32 / pth_gae_ragged_f32)
  ->  PPO.train (pth_index_build + pth_perm_feistel + pth_ppo_update)
which is what ``trainer.py ENV PPO PPO`` executes through
``ego.learn()`` -> ``MultiAgentEnv.step`` -> ``OnPolicyAgent.get_action/update``
(pantheonrl/common/agents.py:111-203), generalised to N envs.
Python only sequences kernel launches; every number is computed on the device.
"""
import dataclasses

import torch

from . import _lib, ops, policy, rollout as ro, update as up


@dataclasses.dataclass
class PPOConfig:
    """SB3 PPO keyword names (trainer.py forwards --ego-config / --alt-config JSON
    verbatim); ``n_minibatches`` replaces ``batch_size`` when set, because SB3's
    default 64 would mean millions of serial Adam steps at N >> 1."""
    learning_rate: float = 3e-4
    n_steps: int = 2048
    batch_size: int = 64
    n_epochs: int = 10
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_range: float = 0.2
    normalize_advantage: bool = True
    ent_coef: float = 0.0
    vf_coef: float = 0.5
    max_grad_norm: float = 0.5
    n_minibatches: int = 0


def split_update_grids(cap, ego_tiles, alt_tiles, floor=48):
    """(ego CTAs, partner CTAs) for two update kernels sharing ``cap`` co-resident CTAs.
    The ego's share is a function of (cap, ego_tiles) ONLY — see VecTrainer.plan_grids."""
    even = lambda t, budget: -(-t // -(-t // budget))  # noqa: E731  t tiles in equal rounds within budget
    ge = max(even(max(ego_tiles, 1), cap // 2), min(cap // 2, floor))
    if alt_tiles <= 0:
        return ge, 0
    return ge, max(even(alt_tiles, cap - ge), min(cap - ge, floor))


class Learner:
    """Device state of one PPO learner: parameters, Adam moments, step counters."""

    def __init__(self, space, cfg, seed, device):
        self.space, self.cfg, self.device = space, cfg, device
        flat = policy.init_flat(space, seed)
        self.params = torch.from_numpy(flat).to(device)
        self.adam_m = torch.zeros_like(self.params)
        self.adam_v = torch.zeros_like(self.params)
        self.adam_step = 0
        self.n_updates = 0
        self.last_stats = None

    def batch_size_for(self, M):
        if self.cfg.n_minibatches > 0:
            return max(1, -(-M // self.cfg.n_minibatches))
        return self.cfg.batch_size

    def state_dict(self):
        return policy.flat_to_state_dict(self.space, self.params.cpu().numpy())


class VecTrainer:
    def __init__(self, env_kind, n_envs, ego_cfg=None, alt_cfg=None, seed=10, partner="ppo",
                 probegostart=0.5, device="cuda", env0=0, group=None, exchange="nccl",
                 ego_update="sharded", layout="simple", ego_agent_idx=0, horizon=None,
                 concurrent_updates=True):
        """group: a torch.distributed process group for one-partner-per-GPU sharding
        (SURVEY.md 8e): every rank owns n_envs envs and its own partner, the ego is
        replicated; per rollout the ranks all-gather their packed ego transitions and
        each runs the same deterministic ego update on the full batch.
        exchange: "nccl" (pack kernel + ncclAllGather) or "p2p" (one kernel packs and
        stores into every rank's gather buffer through NVLink peer mappings).
        ego_update: "sharded" — rank r computes every world-th tile of each global
        minibatch and the per-rank gradient sums are exchanged through peer memory inside
        the update kernel (added in rank order: replicas stay bit-identical);
        "replicated" — every rank computes the whole update redundantly.
        concurrent_updates: run the partner's update kernel next to the ego's (see train())."""
        if not torch.cuda.is_available():
            raise _lib.PthError("VecTrainer needs a CUDA device: the hot path has no CPU implementation")
        self.env_kind, self.N, self.seed = env_kind, int(n_envs), int(seed)
        self.device, self.env0, self.probegostart = device, env0, probegostart
        self.concurrent_updates = bool(concurrent_updates)
        self._side_stream = torch.cuda.Stream(device=device)
        self.last_grids = (0, 0)
        self.space = ro.space_for(env_kind)
        self.d_layout, box, state_bytes = None, False, 32
        if env_kind == "overcooked":  # config 4: OvercookedMultiEnv-v0, Box(62) observations
            from .envs import overcooked as oc
            self.layout = oc.build_layout(layout, ego_agent_idx, horizon or oc.HORIZON)
            self.d_layout = oc.layout_to_device(self.layout, device)
            box, state_bytes = True, _lib.PTH_OC_STATE_BYTES
        self.ego_cfg = ego_cfg or PPOConfig()
        self.alt_cfg = alt_cfg or self.ego_cfg
        self.partner = partner
        self.T = self.ego_cfg.n_steps
        self.ego = Learner(self.space, self.ego_cfg, seed, device)
        if partner == "ppo":
            self.alt = Learner(self.space, self.alt_cfg, seed, device)  # same seed: identical init
        elif partner == "selfplay":
            self.alt = None  # StaticPolicyAgent(ego.policy): shares the ego's weights
        else:
            raise ValueError(partner)
        N, T = self.N, self.T
        # exchange record: observation row | action 4 | logp 4 | advantage 4 | return 4
        self.obs_bytes = 4 * _lib.PTH_OC_ROW if box else 32
        self.rec_bytes = self.obs_bytes + 16
        self.ego_buf = ro.Buffer(T, N, False, device, box)
        alt_cap = ro.alt_capacity(env_kind, T)
        self.alt_buf = ro.Buffer(alt_cap, N, True, device, box)
        self.carry = ro.Carry(N, device, state_bytes)
        self.rollouts = 0
        self.tick_base = 0  # global tick of this trainer's first rollout (a resumed model starts past its old ticks)
        self.num_timesteps = 0
        self.partner_decisions = 0
        # dense env-major sample index of the ego buffer (SB3 swap_and_flatten), built once
        self.group, self.exchange = group, exchange
        self.world = 1
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world == 1:
            self.ego_index, _ = up.index_build(None, T, N, device=device)
        else:
            from .dist_util import global_env_major_index
            self.ego_index = global_env_major_index(self.world, T, N, device)
            self._setup_exchange()
            self.peers = up.PeerExchange(self.space, group, device) if ego_update == "sharded" else None
        self.ego_M = self.world * T * N
        self.ego_perm = torch.empty(self.ego_cfg.n_epochs, self.ego_M, dtype=torch.int32, device=device)
        self.ego_ws = up.UpdateWorkspace(self.space, self.ego_M, self.ego.batch_size_for(self.ego_M), device)
        if self.alt is not None:
            cap = alt_cap * N
            self.alt_perm_store = torch.empty(self.alt_cfg.n_epochs * cap, dtype=torch.int32, device=device)
            self.alt_ws = up.UpdateWorkspace(self.space, cap, max(1, self.alt.batch_size_for(cap)), device)

    # ------------------------------------------------------------------ multi-GPU exchange
    def _setup_exchange(self):
        import torch.distributed as dist
        count = self.T * self.N
        nbytes = self.world * count * self.rec_bytes
        lib, ctx = _lib.load(), _lib.Context.get(torch.device(self.device).index or 0)
        if self.exchange == "p2p":
            import torch.distributed._symmetric_memory as symm
            self.gather = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.symm = symm.rendezvous(self.gather, self.group)
            self.peer_ptrs = torch.tensor(list(self.symm.buffer_ptrs), dtype=torch.int64, device=self.device)
        else:
            self.gather = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.packed = torch.empty(count * self.rec_bytes, dtype=torch.uint8, device=self.device)
            if self.exchange == "nccl" and not ctx.has_comm:
                # the library's own NCCL communicator (pth_comm_init): torch.distributed only carries
                # the 128-byte unique id from rank 0 to the others
                import ctypes as C
                uid = (C.c_uint8 * 128)()
                if self.rank == 0:
                    _lib.check(lib.pth_comm_unique_id(uid), "pth_comm_unique_id")
                box = [bytes(uid)]
                dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0), group=self.group)
                uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
                _lib.check(lib.pth_comm_init(ctx.handle, uid, self.world, self.rank), "pth_comm_init")
                ctx.has_comm = True
        dist.barrier(self.group)

    def pack_into(self, out):
        """This rank's ego transitions of the last rollout as exchange records (pth_pack_transitions)."""
        b, count = self.ego_buf, self.T * self.N
        lib, ctx = _lib.load(), _lib.Context.get(torch.device(self.device).index or 0)
        _lib.check(lib.pth_pack_transitions(ctx.handle, b.obs.data_ptr(), self.obs_bytes, b.actions.data_ptr(),
                                            b.logp.data_ptr(), b.advantages.data_ptr(), b.returns.data_ptr(),
                                            count, out.data_ptr(), _lib.current_stream()), "pth_pack_transitions")
        _lib.count_launch()

    def exchange_ego(self):
        """All-gather this rollout's ego transitions (one exchange per rollout)."""
        import torch.distributed as dist
        b, count = self.ego_buf, self.T * self.N
        lib, ctx = _lib.load(), _lib.Context.get(torch.device(self.device).index or 0)
        if self.exchange == "p2p":
            self.symm.barrier()  # peers finished reading the previous rollout's records
            _lib.check(lib.pth_pack_allgather_p2p(ctx.handle, b.obs.data_ptr(), self.obs_bytes, b.actions.data_ptr(),
                                                  b.logp.data_ptr(), b.advantages.data_ptr(),
                                                  b.returns.data_ptr(), count, self.peer_ptrs.data_ptr(),
                                                  self.world, self.rank, _lib.current_stream()),
                       "pth_pack_allgather_p2p")
            _lib.count_launch()
            self.symm.barrier()  # every rank's stores have landed
            return
        self.pack_into(self.packed)
        if self.exchange == "nccl":  # ncclAllGather on the library's communicator, same stream
            _lib.check(lib.pth_allgather_transitions(ctx.handle, self.packed.data_ptr(), self.packed.numel(),
                                                     self.gather.data_ptr(), _lib.current_stream()),
                       "pth_allgather_transitions")
        else:  # "torch": the same collective through torch.distributed
            dist.all_gather_into_tensor(self.gather, self.packed, group=self.group)

    def learners(self):
        """Every learner whose state this trainer owns: the ego first, then the partner(s)."""
        return [self.ego] + ([self.alt] if self.alt is not None else [])

    # ------------------------------------------------------------------ phases
    def collect(self):
        alt_params = self.alt.params if self.alt is not None else self.ego.params
        ro.run_rollout(self.env_kind, self.space, self.ego.params, alt_params, self.ego_buf,
                       self.alt_buf, self.carry, self.T, self.seed, self.tick_base + self.rollouts * self.T,
                       env0=self.env0, probegostart=self.probegostart,
                       first_rollout=self.rollouts == 0, partner_records=self.alt is not None,
                       d_layout=self.d_layout)
        self.rollouts += 1
        self.num_timesteps += self.N * self.T

    def compute_gae(self):
        b, c = self.ego_buf, self.ego_cfg
        ops.gae(b.rewards, b.values, b.episode_starts, self.carry.ego_last_value,
                self.carry.ego_last_done, c.gamma, c.gae_lambda, out=(b.advantages, b.returns))
        if self.alt is not None:
            a, ac = self.alt_buf, self.alt_cfg
            if self.env_kind == "liar":
                ops.gae_ragged(a.rewards, a.values, a.episode_starts, a.count, self.carry.alt_boot_done,
                               ac.gamma, ac.gae_lambda, out=(a.advantages, a.returns))
            else:
                # simultaneous games: one partner row per tick, the buffer is dense -> the dense kernel, with the
                # partner's bootstrap rule (value of its last stored row, agents.py:127-129)
                ops.gae(a.rewards, a.values, a.episode_starts, a.values[self.T - 1], self.carry.alt_boot_done,
                        ac.gamma, ac.gae_lambda, out=(a.advantages, a.returns))

    def plan_grids(self, M_ego, M_alt):
        """CTAs for the two learners' update kernels when they run side by side.

        Each update is one persistent cooperative kernel with one CTA per SM; a minibatch has
        ceil(batch / 128) tiles (a rank of a sharded ego update computes every world-th one).
        The ego gets at most half of the SMs, in as few equal rounds of tiles as that allows;
        the partner gets what is left, the same way.  Spare SMs are handed out too: CTAs without
        a tile still share the ordered reduction and Adam.

        The grid is part of the update's reduction contract and, for a sharded or
        replicated ego update, must be THE SAME ON EVERY RANK: the ego's grid therefore depends
        only on the global ego batch, the world size and the device — never on this rank's
        (ragged) partner batch.  ``last_grids`` records both for whoever replays the update on
        the oracle."""
        cap = up.update_grid(self.space, 1 << 30, 1 << 30, device=torch.device(self.device).index or 0)
        tiles = lambda M, bs, w: -(-(-(-min(bs, M) // 128)) // w)  # noqa: E731
        sharded = self.world > 1 and getattr(self, "peers", None) is not None
        te = tiles(M_ego, self.ego.batch_size_for(M_ego), self.world if sharded else 1)
        ta = tiles(M_alt, self.alt.batch_size_for(M_alt), 1) if M_alt > 0 else 0
        return split_update_grids(cap, te, ta)

    def _train_one(self, learner, buf, index, M, perm, ws, stream_id, packed=None, peers=None, grid=0):
        cfg = learner.cfg
        up.perm_feistel(M, cfg.n_epochs, self.seed, stream_id, epoch0=learner.n_updates, out=perm)
        bs = learner.batch_size_for(M)
        if packed is None:
            arrays, stride = (buf.obs, buf.actions, buf.logp, buf.advantages, buf.returns), 0
        else:  # read the all-gathered 48-byte records in place
            ob = self.obs_bytes
            arrays, stride = tuple(packed[o:] for o in (0, ob, ob + 4, ob + 8, ob + 12)), self.rec_bytes
            if self.ego_buf.box:  # the kernel takes fp32 rows: same bytes, viewed as floats (272 % 4 == 0)
                arrays = (arrays[0].view(torch.float32).view(-1, self.rec_bytes // 4)[:, :_lib.PTH_OC_ROW],) + arrays[1:]
        stats = up.ppo_update(
            learner.space, learner.params, learner.adam_m, learner.adam_v, learner.adam_step,
            *arrays, perm, bs, ws, index=index, rec_stride=stride, peers=peers,
            M=M, learning_rate=cfg.learning_rate, clip_range=cfg.clip_range, ent_coef=cfg.ent_coef,
            vf_coef=cfg.vf_coef, max_grad_norm=cfg.max_grad_norm,
            normalize_advantage=cfg.normalize_advantage, grid_ctas=grid)
        n_mb = -(-M // bs)
        learner.adam_step += cfg.n_epochs * n_mb
        learner.n_updates += cfg.n_epochs
        learner.last_stats = stats
        return stats

    def train(self):
        """PPO.train for both learners.  They are independent (own parameters, own buffers, own
        shuffle streams), so with ``concurrent_updates`` the partner's update kernel runs on a
        side stream next to the ego's, each on its share of the SMs (plan_grids)."""
        peers = getattr(self, "peers", None) if self.world > 1 else None
        M = 0
        if self.alt is not None:
            a = self.alt_buf
            index, total = up.index_build(a.count, a.Tcap, self.N, device=self.device)
            M = int(total.item())  # the one host read-back per train(): ragged sample count
            self.partner_decisions += M
        both = self.alt is not None and M > 0
        planned = self.alt is not None and self.concurrent_updates  # rank-independent (M is not)
        g_ego, g_alt = self.plan_grids(self.ego_M, M) if planned else (0, 0)
        self.last_grids = (g_ego, g_alt)
        cur = torch.cuda.current_stream()
        if both:
            perm = self.alt_perm_store[: self.alt_cfg.n_epochs * M].view(self.alt_cfg.n_epochs, M)
            side = self._side_stream if self.concurrent_updates else cur
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._train_one(self.alt, a, index, M, perm, self.alt_ws, _lib.STREAM_SHUFFLE_ALT, grid=g_alt)
        packed = None
        if self.world > 1:
            # after the partner's launch: its update needs nothing from the other ranks and runs
            # on its share of the SMs while the ego transitions are exchanged
            self.exchange_ego()
            packed = self.gather
        self._train_one(self.ego, self.ego_buf, self.ego_index, self.ego_M, self.ego_perm,
                        self.ego_ws, _lib.STREAM_SHUFFLE_EGO, packed=packed, peers=peers, grid=g_ego)
        if both:
            cur.wait_stream(side)
        return M

    def iteration(self):
        """collect -> GAE -> train for both agents. Returns agent decisions made."""
        self.collect()
        self.compute_gae()
        m_alt = self.train()
        if self.alt is None:
            m_alt = self.N * self.T  # the static partner still decides once per tick (RPS)
        return self.N * self.T + m_alt

    def learn(self, total_timesteps):
        """SB3-style: run iterations until the ego has taken total_timesteps steps."""
        while self.num_timesteps < total_timesteps:
            self.iteration()
        return self

    # ------------------------------------------------------------------ recording
    def recorded_transitions(self, env=0):
        """What `recorder_wrap(env)` (pantheonrl/common/wrappers.py:84-232) would hold for env `env` after
        the LAST rollout, cut out of the rollout buffers on the host (pantheonrl_b200/vec_record.py; pinned
        against the reference's recorder through the oracle's buffers in tests/test_vec_record_cpu.py).
        Call it between collect() and the next collect()."""
        from . import vec_record as vr
        if self.alt is None:
            raise _lib.PthError("recording needs a recording partner (partner='ppo'): a static partner stores no rows")
        host = lambda b: {k: getattr(b, k).cpu().numpy() for k in ("obs", "actions", "episode_starts")}  # noqa: E731
        ego, alt = host(self.ego_buf), host(self.alt_buf)
        # every recorded partner decision: the complete rows plus the one still open at the rollout's end
        alt["count"] = self.alt_buf.count.cpu().numpy() + ((self.carry.flags.cpu().numpy() >> 2) & 1)
        last_done = float(self.carry.ego_last_done[env].item())
        if self.env_kind == "liar":
            if self.rollouts != 1:
                # a later rollout may begin after the partner's opening move of the running episode, which sits
                # in the PREVIOUS partner buffer: the episode counters of the two buffers then differ by one
                raise _lib.PthError("turn-based recording is cut from a rollout that starts at an episode "
                                    "boundary: call it after the first collect()")
            return vr.turn_based_transitions(ego, alt, env, last_done)
        obs_len, act_len = (1, 1) if self.env_kind == "rps" else (self.space.obs_len, 1)
        return vr.simultaneous_transitions(ego, alt, env, last_done, obs_len=obs_len, act_len=act_len)

    # ------------------------------------------------------------------ logging
    def train_stats(self, learner=None):
        """Mean of the per-minibatch scalars SB3's PPO.train logs (adap_learn.py:354-371)."""
        learner = learner or self.ego
        s = learner.last_stats.cpu().numpy()
        return {"train/policy_gradient_loss": float(s[:, 0].mean()), "train/value_loss": float(s[:, 1].mean()),
                "train/entropy_loss": float(s[:, 2].mean()),
                "train/approx_kl": float(s[-max(1, len(s) // max(1, learner.cfg.n_epochs)):, 3].mean()),  # last epoch (SB3)
                "train/clip_fraction": float(s[:, 4].mean()), "train/loss": float(s[-1, 5]),
                "train/n_updates": learner.n_updates}

    def episode_stats(self):
        e = self.carry.ep_stats.cpu().numpy()
        eps = max(1.0, float(e[0]))
        return {"rollout/ep_rew_mean": float(e[1]) / eps, "rollout/ep_len_mean": float(e[2]) / eps,
                "episodes": int(e[0]), "ego_steps": int(e[2]), "partner_decisions": int(e[3])}
