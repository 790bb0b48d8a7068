"""pantheonrl_b200 — B200-native PPO rollout / GAE / update engine behind the
PantheonRL agent + multi-agent-env plugin surface.

The hot path lives in ``libpantheon_b200.so`` (hand-written sm_100a CUDA behind
the C ABI of ``include/pantheon_b200.h``); this package is the thin host side.
"""
__version__ = "0.1.0"
