"""GPU parity: the game-rule kernels vs traces of the reference's own classes
(tests/golden) and vs the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import oracle
from pantheonrl_b200 import ops
from test_oracle_cpu import replay_liar_env, _state_from_hands

pytestmark = pytest.mark.gpu


def test_rps_payoff_matches_reference(ctx, golden_dir):
    tab = np.load(os.path.join(golden_dir, "rps_payoff.npz"))["table"]
    re, ra = ops.rps_step(torch.tensor(tab[:, 0], dtype=torch.int32).cuda(),
                          torch.tensor(tab[:, 1], dtype=torch.int32).cuda())
    assert np.array_equal(re.cpu().numpy(), tab[:, 2].astype(np.float32))
    assert np.array_equal(ra.cpu().numpy(), tab[:, 3].astype(np.float32))


def _gpu_liar_step(state, is_ego, action):
    st = torch.from_numpy(state).cuda()
    obs, re, ra, done = ops.liar_step(st, torch.from_numpy(np.ascontiguousarray(is_ego, np.uint8)).cuda(),
                                      torch.from_numpy(np.ascontiguousarray(action, np.uint8)).cuda())
    state[:] = st.cpu().numpy()
    return obs.cpu().numpy(), re.cpu().numpy(), ra.cpu().numpy(), done.cpu().numpy()


def test_liar_env_matches_reference_trace(ctx, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "liar_env.npz")))
    replay_liar_env(g, _gpu_liar_step)


def test_liar_random_play_matches_oracle(ctx):
    N = 50000
    rng = np.random.RandomState(0)
    s_gpu, ef_gpu, obs_gpu = ops.liar_reset(N, seed=10, tick=5, env0=123)
    s_cpu, ef_cpu, obs_cpu = oracle.liar_reset(N, seed=10, tick=5, env0=123)
    assert np.array_equal(s_gpu.cpu().numpy(), s_cpu)
    assert np.array_equal(ef_gpu.cpu().numpy(), ef_cpu)
    assert np.array_equal(obs_gpu.cpu().numpy(), obs_cpu)
    turn = ef_cpu.copy()
    for k in range(14):
        act = np.stack([rng.randint(7, size=N), rng.randint(12, size=N)], 1).astype(np.uint8)
        if k < 8:
            act[:, 1] = np.minimum(11, s_cpu[:, 24] + (rng.rand(N) < 0.9))  # keep games alive a while
            act[:, 0] = rng.randint(6, size=N)
        o0, re0, ra0, d0 = oracle.liar_step(s_cpu, turn, act)
        o1, re1, ra1, d1 = ops.liar_step(s_gpu, torch.from_numpy(turn).cuda(), torch.from_numpy(act).cuda())
        assert np.array_equal(o1.cpu().numpy(), o0)
        assert np.array_equal(re1.cpu().numpy(), re0) and np.array_equal(ra1.cpu().numpy(), ra0)
        assert np.array_equal(d1.cpu().numpy(), d0)
        assert np.array_equal(s_gpu.cpu().numpy(), s_cpu)
        turn = 1 - turn


def test_liar_forced_bluff_after_twelve_bids(ctx):
    hands = np.array([[1, 1, 1, 1, 1, 1, 6, 0, 0, 0, 0, 0]], np.uint8)
    st = torch.from_numpy(_state_from_hands(hands)).cuda()
    ego = 1
    for c in range(12):
        obs, re, ra, d = ops.liar_step(st, torch.tensor([ego], dtype=torch.uint8).cuda(),
                                       torch.tensor([[0, c]], dtype=torch.uint8).cuda())
        assert d.item() == 0
        ego ^= 1
    obs, re, ra, d = ops.liar_step(st, torch.tensor([ego], dtype=torch.uint8).cuda(),
                                   torch.tensor([[3, 11]], dtype=torch.uint8).cuda())
    assert d.item() == 1 and re.item() == 1.0 and ra.item() == -1.0
