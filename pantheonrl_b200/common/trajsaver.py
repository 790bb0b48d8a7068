"""Recorded multi-agent transitions and their ``.npy`` file format.

API and on-disk layout of pantheonrl/common/trajsaver.py: ``TransitionsMinimal``
(:42-131), ``TurnBasedTransitions`` (:145-177), ``SimultaneousTransitions``
(:180-232).  A file is ONE 2-D array, one row per recorded step:

  TransitionsMinimal       [ obs | acts ]
  TurnBasedTransitions     [ obs | acts | flag ]                     flag: who moved, and whether the game ended
  SimultaneousTransitions  [ egoobs | egoacts | altobs | altacts | flag ]

with every field flattened per row; readers split the columns by the sizes of the
spaces.  Host-side data plumbing (numpy); consumed by behaviour cloning in the
reference (pantheonrl/algos/bc.py:270-363).
"""
import dataclasses

import numpy as np

from ..spaces import Box, Discrete, MultiDiscrete


def get_space_size(space):
    """Columns one observation / action occupies in a file (util.py:14-29)."""
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return 1
    if isinstance(space, MultiDiscrete):
        return len(space.nvec)
    raise ValueError(f"unsupported space {space!r}")


def _rows(a, n):
    return np.reshape(a, (n, -1))


@dataclasses.dataclass(frozen=True)
class TransitionsMinimal:
    """obs[i] is what the agent saw when it chose acts[i]; usable as a torch ``Dataset``
    (``len``, integer index -> dict of arrays, slice -> TransitionsMinimal)."""
    obs: np.ndarray
    acts: np.ndarray

    def __post_init__(self):
        for v in (self.obs, self.acts):
            if isinstance(v, np.ndarray):
                v.setflags(write=False)
        if len(self.obs) != len(self.acts):
            raise ValueError(f"obs and acts must have same number of timesteps: {len(self.obs)} != {len(self.acts)}")

    def __len__(self):
        return len(self.obs)

    def __getitem__(self, key):
        if isinstance(key, slice):
            return TransitionsMinimal(self.obs[key], self.acts[key])
        return {"obs": self.obs[key], "acts": self.acts[key]}

    def write_transition(self, file):
        np.save(file, np.concatenate((self.obs, self.acts), axis=1))

    @classmethod
    def read_transition(cls, file, obs_space, act_space):
        table = np.load(file)
        k = get_space_size(obs_space)
        return cls(table[:, :k], table[:, k:])


class MultiTransitions:
    """Both players' transitions of one recording."""

    def get_ego_transitions(self):
        raise NotImplementedError

    def get_alt_transitions(self):
        raise NotImplementedError


@dataclasses.dataclass(frozen=True)
class TurnBasedTransitions(MultiTransitions):
    """One row per move; flags: 0 ego moved, 1 partner moved, 2 / 3 the same and the game ended."""
    obs: np.ndarray
    acts: np.ndarray
    flags: np.ndarray

    def _side(self, parity):
        mask = np.asarray(self.flags) % 2 == parity
        return TransitionsMinimal(self.obs[mask], self.acts[mask])

    def get_ego_transitions(self):
        return self._side(0)

    def get_alt_transitions(self):
        return self._side(1)

    def write_transition(self, file):
        n = np.size(self.flags)
        np.save(file, np.concatenate((_rows(self.obs, n), _rows(self.acts, n), _rows(self.flags, n)), axis=1))

    @classmethod
    def read_transition(cls, file, obs_space, act_space):
        table = np.load(file)
        k = get_space_size(obs_space)
        return cls(table[:, :k], table[:, k:-1], table[:, -1])


@dataclasses.dataclass(frozen=True)
class SimultaneousTransitions(MultiTransitions):
    """One row per joint step; flags: 0 game continues, 1 game ended."""
    egoobs: np.ndarray
    egoacts: np.ndarray
    altobs: np.ndarray
    altacts: np.ndarray
    flags: np.ndarray

    def get_ego_transitions(self):
        return TransitionsMinimal(self.egoobs, self.egoacts)

    def get_alt_transitions(self):
        return TransitionsMinimal(self.altobs, self.altacts)

    def write_transition(self, file):
        n = np.size(self.flags)
        cols = (self.egoobs, self.egoacts, self.altobs, self.altacts, self.flags)
        np.save(file, np.concatenate([_rows(c, n) for c in cols], axis=1))

    @classmethod
    def read_transition(cls, file, obs_space, act_space):
        table = np.load(file)
        k, m = get_space_size(obs_space), get_space_size(act_space)
        return cls(table[:, :k], table[:, k:k + m], table[:, k + m:2 * k + m], table[:, 2 * k + m:-1], table[:, -1])
