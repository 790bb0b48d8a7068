"""CPU ORACLE (test infrastructure): literal numpy restatement of
stable-baselines3 1.7.0 RolloutBuffer.compute_returns_and_advantage — the loop
over reversed(range(buffer_size)) with float32 numpy vectors of length n_envs —
as called from pantheonrl/common/agents.py:127-130.  Same recurrence in-tree:
overcookedgym/human_aware_rl/baselines/baselines/ppo2/runner.py:152-164.
SB3 is not vendored; the in-tree loop is executed verbatim by
tests/golden/make_golden_sb3_intree.py and this file / pth_oracle.c reproduce its
outputs to 1e-5 (tests/test_oracle_sb3_intree.py; the in-tree loop accumulates in
float64, SB3's partner path and this file in float32).
"""
import numpy as np


def compute_returns_and_advantage(rewards, values, episode_starts, last_values, dones,
                                  gamma=0.99, gae_lambda=0.95):
    rewards = np.asarray(rewards, np.float32)
    values = np.asarray(values, np.float32)
    episode_starts = np.asarray(episode_starts, np.float32)
    last_values = np.asarray(last_values, np.float32).flatten()
    dones = np.asarray(dones, np.float32)
    T = rewards.shape[0]
    advantages = np.zeros_like(rewards)
    last_gae_lam = 0
    for step in reversed(range(T)):
        if step == T - 1:
            next_non_terminal = 1.0 - dones
            next_values = last_values
        else:
            next_non_terminal = 1.0 - episode_starts[step + 1]
            next_values = values[step + 1]
        delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
        last_gae_lam = delta + gamma * gae_lambda * next_non_terminal * last_gae_lam
        advantages[step] = last_gae_lam
    returns = advantages + values
    return advantages, returns
