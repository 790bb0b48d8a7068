"""GPU parity: the rollout megakernel vs the CPU oracle — every buffer the two
learners would train on, bit for bit (action / reward / done traces, values,
log-probs, episode_start flags, ragged partner cursors, carried env state)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import rollout as orc
from pantheonrl_b200 import rollout as dev
from test_oracle_cpu import rand_params

pytestmark = pytest.mark.gpu

KW = {"rps": oracle.RPS_SPACE, "liar": oracle.LIAR_SPACE}


def _gpu_rollout(env_kind, pe, pa, N, T, seed, tick0=0, env0=0, first=True, carry=None, selfplay=False,
                 records=True, probegostart=0.5, alt=None):
    sp = dev.space_for(env_kind)
    d_pe = torch.from_numpy(pe).cuda()
    d_pa = d_pe if selfplay else torch.from_numpy(pa).cuda()
    ego = dev.Buffer(T, N, False, "cuda")
    alt = alt or dev.Buffer(dev.alt_capacity(env_kind, T), N, True, "cuda")
    carry = carry or dev.Carry(N, "cuda")
    dev.run_rollout(env_kind, sp, d_pe, d_pa, ego, alt, carry, T, seed, tick0, env0=env0,
                    first_rollout=first, partner_records=records, probegostart=probegostart)
    torch.cuda.synchronize()
    return ego, alt, carry


def _compare(ego, alt, carry, o_ego, o_alt, o_carry, records=True):
    for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
        assert np.array_equal(getattr(ego, k).cpu().numpy(), o_ego[k]), f"ego {k}"
    for k in ("ego_last_start", "alt_last_done", "total_rew", "flags", "ego_last_value", "ego_last_done",
              "alt_boot_done"):
        assert np.array_equal(getattr(carry, k).cpu().numpy(), o_carry[k]), f"carry {k}"
    assert np.array_equal(carry.game_state.cpu().numpy()[:, :25], o_carry["game_state"][:, :25])
    assert np.array_equal(carry.ep_stats.cpu().numpy(), o_carry["ep_stats"])
    if records:
        cnt = alt.count.cpu().numpy()
        assert np.array_equal(cnt, o_alt["count"])
        Tc = min(alt.Tcap, o_alt["obs"].shape[0])
        # complete rows + the row left open for the next rollout (flags bit2), which sits at row count[n]
        mask = np.arange(Tc)[:, None] < (cnt + ((o_carry["flags"] >> 2) & 1))[None, :]
        for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
            g = getattr(alt, k).cpu().numpy()[:Tc]
            w = o_alt[k][:Tc]
            m = mask.reshape(mask.shape + (1,) * (g.ndim - 2))
            assert np.array_equal(np.where(m, g, 0), np.where(m, w, 0)), f"alt {k}"


@pytest.mark.parametrize("env_kind,N,T", [("rps", 300, 16), ("liar", 128, 32), ("liar", 1000, 24), ("liar", 1, 64)])
def test_rollout_bit_exact_vs_oracle(ctx, env_kind, N, T):
    osp = oracle.make_space(**KW[env_kind])
    pe = rand_params(osp, seed=1, scale=0.3)
    pa = rand_params(osp, seed=2, scale=0.3)
    o = orc.rollout(env_kind, osp, pe, pa, N=N, T=T, seed=10, tick0=5, env0=7,
                    alt=orc.new_buffer(orc.alt_capacity(env_kind, T), N, True))
    g = _gpu_rollout(env_kind, pe, pa, N, T, seed=10, tick0=5, env0=7)
    _compare(*g, *o)


def test_rollout_carry_across_rollouts(ctx):
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    pe = rand_params(osp, seed=3, scale=0.2)
    pa = rand_params(osp, seed=4, scale=0.2)
    N, T = 500, 16
    o1 = orc.rollout("liar", osp, pe, pa, N=N, T=T, seed=2)
    g1 = _gpu_rollout("liar", pe, pa, N, T, seed=2)
    _compare(*g1, *o1)
    open1 = (o1[2]["flags"] >> 2) & 1
    assert 0 < open1.sum() < N  # some partner rows wait for the ego's next move: carried, not dropped
    # the same partner buffer goes into the next rollout (the open row is in it, at row count[n])
    o2 = orc.rollout("liar", osp, pe, pa, N=N, T=T, seed=2, tick0=T, first_rollout=False, carry=o1[2], alt=o1[1])
    g2 = _gpu_rollout("liar", pe, pa, N, T, seed=2, tick0=T, first=False, carry=g1[2], alt=g1[1])
    _compare(*g2, *o2)
    # the boundary case of the partner's reward: a carried row finished by the ego's first move
    r0 = g2[1].rewards[0].cpu().numpy()
    assert np.any((open1 == 1) & (r0 != 0))
    o3 = orc.rollout("liar", osp, pe, pa, N=N, T=T, seed=2, tick0=2 * T, first_rollout=False, carry=o2[2], alt=o2[1])
    g3 = _gpu_rollout("liar", pe, pa, N, T, seed=2, tick0=2 * T, first=False, carry=g2[2], alt=g2[1])
    _compare(*g3, *o3)


def test_selfplay_static_partner(ctx):
    # BASELINE config 3 shape: partner = StaticPolicyAgent(ego.policy) (agents.py:54-79):
    # same weights, nothing recorded for the partner.
    osp = oracle.make_space(**oracle.RPS_SPACE)
    pe = rand_params(osp, seed=5, scale=0.8)
    N, T = 4096, 8
    o = orc.rollout("rps", osp, pe, pe, N=N, T=T, seed=1, partner_records=False,
                    alt=orc.new_buffer(T, N, True))
    g = _gpu_rollout("rps", pe, pe, N, T, seed=1, selfplay=True, records=False)
    _compare(*g, *o, records=False)
    r = g[0].rewards
    assert set(r.unique().tolist()) <= {-1.0, 0.0, 1.0}


def test_sharded_rollout_equals_global(ctx):
    # multi-GPU sharding contract: env0 offsets index one global RNG stream
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    pe = rand_params(osp, seed=6, scale=0.3)
    pa = rand_params(osp, seed=7, scale=0.3)
    T = 12
    full = _gpu_rollout("liar", pe, pa, 512, T, seed=4)
    lo = _gpu_rollout("liar", pe, pa, 256, T, seed=4, env0=0)
    hi = _gpu_rollout("liar", pe, pa, 256, T, seed=4, env0=256)
    for k in ("obs", "actions", "rewards", "values", "logp"):
        f = getattr(full[0], k)
        assert torch.equal(f[:, :256], getattr(lo[0], k)) and torch.equal(f[:, 256:], getattr(hi[0], k))


def test_full_size_rollout_properties(ctx):
    # BASELINE config 2 size (4096 envs, T = 128): properties + sampled oracle check
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    pe = rand_params(osp, seed=8, scale=0.1)
    pa = rand_params(osp, seed=9, scale=0.1)
    N, T = 4096, 128
    ego, alt, carry = _gpu_rollout("liar", pe, pa, N, T, seed=10)
    r = ego.rewards
    assert set(r.unique().tolist()) <= {-1.0, 0.0, 1.0}
    nxt = torch.cat([ego.episode_starts[1:], carry.ego_last_done[None]])
    assert bool(((r != 0) <= (nxt == 1)).all())
    cnt = alt.count
    assert int(cnt.min()) >= 1 and int(cnt.max()) <= 2 * T
    st = carry.ep_stats.cpu().numpy()
    n_open = int(((carry.flags >> 2) & 1).sum())  # partner rows still waiting for the ego's next move
    assert 0 < n_open < N
    assert st[2] == N * T and st[3] == int(cnt.sum()) + n_open and st[0] == int((nxt == 1).sum())
    # zero-sum up to the rewards on rows still open at the end (carried into the next rollout)
    mask = torch.arange(alt.Tcap, device="cuda")[:, None] < cnt[None, :]
    assert abs(float(r.sum()) + float((alt.rewards * mask).sum())) <= N
    # a strided sample of envs replayed on the oracle
    idx = np.arange(0, N, 257)
    for n in idx:
        o = orc.rollout("liar", osp, pe, pa, N=1, T=T, seed=10, env0=int(n))
        assert np.array_equal(ego.actions[:, n].cpu().numpy(), o[0]["actions"][:, 0])
        assert np.array_equal(ego.rewards[:, n].cpu().numpy(), o[0]["rewards"][:, 0])
        assert np.array_equal(ego.values[:, n].cpu().numpy(), o[0]["values"][:, 0])
