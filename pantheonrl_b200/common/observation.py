"""Observation container handed to agents (API of pantheonrl/common/observation.py:7-33)."""
import numpy as np


class Observation:
    """obs: what the agent sees; state: full state (defaults to obs); action_mask: legal actions or None."""

    __slots__ = ("obs", "state", "action_mask")

    def __init__(self, obs, state=None, action_mask=None):
        self.obs = obs
        self.state = obs if state is None else state
        self.action_mask = action_mask

    def __repr__(self):
        return f"Observation(obs={np.asarray(self.obs).tolist()!r})"


def extract_obs(observation):
    """Default ego extractor: SB3-style learners only want the array."""
    return observation.obs


def extract_partial_obs(observation):
    return observation.obs, observation.action_mask
