"""CPU ORACLE (test infrastructure): literal numpy restatement of
stable-baselines3 1.7.0 RolloutBuffer.compute_returns_and_advantage — the loop
over reversed(range(buffer_size)) with float32 numpy vectors of length n_envs —
as called from pantheonrl/common/agents.py:127-130.  Same recurrence in-tree:
overcookedgym/human_aware_rl/baselines/baselines/ppo2/runner.py:150-165.
"parity unpinned": SB3 is not vendored and the reference has no golden vectors
for it; this file and pth_oracle.c are independent restatements of each other.
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
