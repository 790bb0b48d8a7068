"""CPU ORACLE (test infrastructure / CPU baseline arm): one full train
iteration of the hot path on the host cores — the same collect -> GAE -> PPO.train
sequence VecTrainer runs on the GPU — composed from the oracle pieces:

  rollout : pth_oracle_rollout.inc (C restatement of MultiAgentEnv.step/reset +
            OnPolicyAgent.get_action/update + the game rules), OpenMP over envs
  GAE     : orc_gae / orc_gae_ragged (SB3 compute_returns_and_advantage)
  update  : oracle/sb3_torch.py — torch CPU eager ops in SB3 PPO.train's order
            (what the reference actually executes on a CPU: torch eager autograd
            + torch.optim.Adam), all host threads

Used by tests (end-to-end cross-check of the engine) and by bench.py's
cpu_baseline / --impl reference legs.  N >> 1 is "not a reference capability":
the reference runs one env; this is its algorithm batched over envs.
"""
import time

import numpy as np
import torch

import oracle
from oracle import rollout as orc
from oracle import sb3_torch
from oracle import update as oupd

STREAM_SHUFFLE_EGO, STREAM_SHUFFLE_ALT = 4, 5


class CpuTrainer:
    def __init__(self, env_kind, n_envs, n_steps=128, n_epochs=10, n_minibatches=32, seed=10,
                 partner="ppo", batch_size=0, layout="simple"):
        """n_minibatches > 0: batch = ceil(M / n_minibatches) (the engine's rule); else SB3's batch_size."""
        self.env_kind, self.N, self.T, self.seed = env_kind, n_envs, n_steps, seed
        self.oc_layout = None
        if env_kind == "overcooked":
            import os
            from oracle import overcooked as ooc
            golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
            self.oc_layout = ooc.multienv_layout(golden, layout)
            self.kw = dict(box_dim=62, heads=[6])
            self.space = oracle.make_space(box_dim=62, heads=[6])
            mk = lambda: sb3_torch.MlpPolicy(box_dim=62, heads=[6], seed=seed)  # noqa: E731
            self.nslot, self.nh, self.row = 62, 1, 64
        else:
            self.kw = oracle.LIAR_SPACE if env_kind == "liar" else oracle.RPS_SPACE
            self.space = oracle.make_space(**self.kw)
            mk = lambda: sb3_torch.MlpPolicy(nvec=self.kw["nvec"], heads=self.kw["heads"], seed=seed)  # noqa: E731
            self.nslot, self.nh, self.row = len(self.kw["nvec"]), len(self.kw["heads"]), 32
        self.n_epochs, self.n_mb, self.batch_size = n_epochs, n_minibatches, batch_size
        self.ego = mk()
        self.alt = mk() if partner == "ppo" else None
        self.carry = None
        # one partner buffer for the whole run: a row left open at a rollout's end is carried in it
        self._alt_buf = orc.new_buffer(orc.alt_capacity(env_kind, n_steps), n_envs, True, env_kind == "overcooked")
        self.rollouts = 0
        self.n_updates = [0, 0]
        self.timing = {}

    def iteration(self):
        N, T = self.N, self.T
        t0 = time.perf_counter()
        pe = self.ego.to_flat()
        pa = self.alt.to_flat() if self.alt is not None else pe
        ego, alt, self.carry = orc.rollout(
            self.env_kind, self.space, pe, pa, N=N, T=T, seed=self.seed, tick0=self.rollouts * T,
            first_rollout=self.rollouts == 0, carry=self.carry, partner_records=self.alt is not None,
            alt=self._alt_buf, oc_layout=self.oc_layout)
        self.rollouts += 1
        t1 = time.perf_counter()
        adv, ret = oracle.gae(ego["rewards"], ego["values"], ego["episode_starts"],
                              self.carry["ego_last_value"], self.carry["ego_last_done"])
        if self.alt is not None:
            aadv, aret = oracle.gae_ragged(alt["rewards"], alt["values"], alt["episode_starts"],
                                           alt["count"], self.carry["alt_boot_done"])
        t2 = time.perf_counter()
        decisions = N * T
        idx = oupd.index_build(None, T, N)
        self._train(self.ego, 0, ego, adv, ret, idx, STREAM_SHUFFLE_EGO)
        if self.alt is not None:
            aidx = oupd.index_build(alt["count"], alt["obs"].shape[0], N)
            decisions += aidx.size
            if aidx.size:
                self._train(self.alt, 1, alt, aadv, aret, aidx, STREAM_SHUFFLE_ALT)
        else:
            decisions += N * T
        t3 = time.perf_counter()
        self.timing = dict(rollout_s=t1 - t0, gae_s=t2 - t1, train_s=t3 - t2)
        return decisions

    def _train(self, pol, which, buf, adv, ret, index, stream):
        M = index.size
        perm = oupd.perm_feistel(M, self.n_epochs, self.seed, stream, epoch0=self.n_updates[which])
        bs = max(1, -(-M // self.n_mb)) if self.n_mb > 0 else self.batch_size
        obs = buf["obs"].reshape(-1, self.row)[index][:, :self.nslot]
        act = buf["actions"].reshape(-1, 4)[index][:, :self.nh]
        sb3_torch.ppo_train(pol, obs, act, buf["logp"].reshape(-1)[index], adv.reshape(-1)[index],
                            ret.reshape(-1)[index], perm, bs)
        self.n_updates[which] += self.n_epochs
