"""Several partner learners per GPU: the reference's partner SET (`env.add_partner_agent` called
more than once, trainer.py:216-228; one partner drawn per episode, multiagentenv.py:113-125, 224)
on the device engine.

The reference draws the partner again at every reset.  The device engine pairs env instances with
partners for good instead ("N/G env instances permanently paired with partner g —
distributionally equivalent to per-episode round-robin"): the N envs of a GPU are cut into P
contiguous groups of N / P, group p always plays partner p.  Every partner sees the same share of
episodes as under round-robin, every ego sample is played against a uniformly chosen partner, and
— what makes it the device-friendly choice — a CTA of the rollout kernel holds ONE partner's
weights in shared memory for its whole horizon, exactly as with a single partner.

  rollout : one launch of pth_rollout_run per group on its own stream (global env ids, so the
            Philox streams are those of one big run); the ego's buffer is laid out
            [P][T][N / P], each group writing its own [T][N / P] block
  GAE     : per group (ego block, partner's ragged buffer)
  update  : ONE ego update over all P blocks (env-major over global env ids; the exchange
            records of a multi-GPU run are the same flat buffer) next to P partner updates,
            every kernel on its planned share of the SMs
With G GPUs, rank r hosts partners r * P .. r * P + P - 1: `bench.py --partners 8 --gpus {1,2,4,8}`
keeps BASELINE configs[4]'s 8 partners and moves them across 1 to 8 GPUs.
"""
import torch

from . import _lib, ops, rollout as ro, update as up
from .engine import Learner, VecTrainer


class _View:
    """Rows [r0, r1) of a Buffer as a Buffer (same storage)."""

    def __init__(self, buf, r0, r1):
        self.Tcap, self.N, self.ragged, self.box, self.count = r1 - r0, buf.N, False, buf.box, None
        for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts", "advantages", "returns"):
            setattr(self, k, getattr(buf, k)[r0:r1])

    c_struct = ro.Buffer.c_struct


class _Lane:
    """One partner and its env group."""

    def __init__(self, tr, p, Ng, alt_cap, state_bytes, ep_stats):
        dev, box = tr.device, tr.ego_buf.box
        self.p, self.Ng = p, Ng
        self.ego_view = _View(tr.ego_buf, p * tr.T, (p + 1) * tr.T)
        self.learner = Learner(tr.space, tr.alt_cfg, tr.seed, dev)  # same seed: identical init (trainer.py:198)
        self.buf = ro.Buffer(alt_cap, Ng, True, dev, box)
        self.carry = ro.Carry(Ng, dev, state_bytes)
        self.carry.ep_stats = ep_stats  # one set of episode counters for the whole trainer
        cap = alt_cap * Ng
        self.perm_store = torch.empty(tr.alt_cfg.n_epochs * cap, dtype=torch.int32, device=dev)
        self.ws = up.UpdateWorkspace(tr.space, cap, max(1, self.learner.batch_size_for(cap)), dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.shuffle_stream = _lib.STREAM_SHUFFLE_ALT if p == 0 else 0x40 + p
        self.M = 0


class PartnerSetTrainer(VecTrainer):
    def __init__(self, env_kind, n_envs, ego_cfg=None, alt_cfg=None, partners_per_gpu=2, **kw):
        P = int(partners_per_gpu)
        if P < 1 or n_envs % P:
            raise ValueError("n_envs must be a multiple of partners_per_gpu")
        if kw.get("partner", "ppo") != "ppo":
            raise ValueError("a partner set is made of learners (partner='ppo')")
        super().__init__(env_kind, n_envs, ego_cfg, alt_cfg, **kw)
        self.P, self.Ng = P, self.N // P
        T, Ng, dev = self.T, self.Ng, self.device
        # the ego's buffer [T][N] re-read as [P][T][Ng]: same bytes, block p belongs to group p
        for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts", "advantages", "returns"):
            t = getattr(self.ego_buf, k)
            setattr(self.ego_buf, k, t.view(P * T, Ng, *t.shape[2:]))
        self.ego_buf.Tcap, self.ego_buf.N = P * T, Ng
        alt_cap = ro.alt_capacity(env_kind, T)
        state_bytes = self.carry.game_state.shape[1]
        ep_stats = self.carry.ep_stats
        self.lanes = [_Lane(self, p, Ng, alt_cap, state_bytes, ep_stats) for p in range(P)]
        # names of the single-partner trainer point at partner 0
        self.alt, self.alt_buf, self.carry = self.lanes[0].learner, self.lanes[0].buf, self.lanes[0].carry
        self.alt_ws = self.alt_perm_store = None
        # env-major sample order over GLOBAL env ids: group p of rank r is "virtual rank" r * P + p
        from .dist_util import global_env_major_index
        self.ego_index = global_env_major_index(self.world * P, T, Ng, dev)
        self.last_grids = (0,) + (0,) * P

    def learners(self):
        return [self.ego] + [ln.learner for ln in self.lanes]

    # ------------------------------------------------------------------ phases
    def collect(self):
        cur = torch.cuda.current_stream()
        tick0 = self.tick_base + self.rollouts * self.T
        for ln in self.lanes:
            ln.stream.wait_stream(cur)
            with torch.cuda.stream(ln.stream):
                ro.run_rollout(self.env_kind, self.space, self.ego.params, ln.learner.params, ln.ego_view, ln.buf,
                               ln.carry, self.T, self.seed, tick0, env0=self.env0 + ln.p * self.Ng,
                               probegostart=self.probegostart, first_rollout=self.rollouts == 0,
                               partner_records=True, d_layout=self.d_layout)
        for ln in self.lanes:
            cur.wait_stream(ln.stream)
        self.rollouts += 1
        self.num_timesteps += self.N * self.T

    def compute_gae(self):
        c, ac = self.ego_cfg, self.alt_cfg
        for ln in self.lanes:
            b, a = ln.ego_view, ln.buf
            ops.gae(b.rewards, b.values, b.episode_starts, ln.carry.ego_last_value, ln.carry.ego_last_done,
                    c.gamma, c.gae_lambda, out=(b.advantages, b.returns))
            ops.gae_ragged(a.rewards, a.values, a.episode_starts, a.count, ln.carry.alt_boot_done,
                           ac.gamma, ac.gae_lambda, out=(a.advantages, a.returns))

    def plan_grids(self, M_ego, M_alts):
        """(ego CTAs, CTAs of partner 0, 1, ...): the ego's share as in the single-partner trainer
        (a function of the global ego batch only: the same on every rank), the partners split the rest."""
        cap = up.update_grid(self.space, 1 << 30, 1 << 30, device=torch.device(self.device).index or 0)
        tiles = lambda M, bs, w: -(-(-(-min(bs, M) // 128)) // w)  # noqa: E731
        even = lambda t, budget: -(-t // -(-t // budget))  # noqa: E731  t tiles in equal rounds within budget
        sharded = self.world > 1 and getattr(self, "peers", None) is not None
        te = tiles(M_ego, self.ego.batch_size_for(M_ego), self.world if sharded else 1)
        ge = max(even(max(te, 1), cap // 2), min(cap // 2, 48))
        share = max(1, (cap - ge) // self.P)
        ga = tuple(even(max(tiles(M, ln.learner.batch_size_for(M), 1), 1), share) if M > 0 else 0
                   for ln, M in zip(self.lanes, M_alts))
        return (ge,) + ga

    def train(self):
        peers = getattr(self, "peers", None) if self.world > 1 else None
        cur = torch.cuda.current_stream()
        built = [up.index_build(ln.buf.count, ln.buf.Tcap, self.Ng, device=self.device) for ln in self.lanes]
        totals = torch.cat([t for _, t in built]).tolist()  # the one host read-back per train()
        grids = self.plan_grids(self.ego_M, totals)
        self.last_grids = grids
        for ln, (index, _), M, g in zip(self.lanes, built, totals, grids[1:]):
            ln.M = int(M)
            if ln.M == 0:
                continue
            perm = ln.perm_store[: self.alt_cfg.n_epochs * ln.M].view(self.alt_cfg.n_epochs, ln.M)
            ln.stream.wait_stream(cur)
            with torch.cuda.stream(ln.stream):
                self._train_one(ln.learner, ln.buf, index, ln.M, perm, ln.ws, ln.shuffle_stream, grid=g)
        packed = None
        if self.world > 1:
            self.exchange_ego()
            packed = self.gather
        self._train_one(self.ego, self.ego_buf, self.ego_index, self.ego_M, self.ego_perm, self.ego_ws,
                        _lib.STREAM_SHUFFLE_EGO, packed=packed, peers=peers, grid=grids[0])
        for ln in self.lanes:
            cur.wait_stream(ln.stream)
        m = int(sum(totals))
        self.partner_decisions += m
        return m

    def recorded_transitions(self, env=0):
        raise _lib.PthError("recorded_transitions: single-partner trainers only")


def make_trainer(env_kind, n_envs, ego_cfg=None, alt_cfg=None, partners_per_gpu=1, **kw):
    """VecTrainer, or PartnerSetTrainer when a GPU hosts more than one partner learner."""
    if partners_per_gpu > 1:
        return PartnerSetTrainer(env_kind, n_envs, ego_cfg, alt_cfg, partners_per_gpu=partners_per_gpu, **kw)
    return VecTrainer(env_kind, n_envs, ego_cfg, alt_cfg, **kw)
