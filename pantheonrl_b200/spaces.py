"""Minimal observation / action spaces (gym is not a dependency).  Attribute
names follow gym.spaces so user code written for the reference keeps working."""
import numpy as np


class Space:
    shape = ()


class Discrete(Space):
    def __init__(self, n):
        self.n, self.shape = int(n), ()

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = (len(self.nvec),)

    def __repr__(self):
        return f"MultiDiscrete({self.nvec.tolist()})"


class Box(Space):
    def __init__(self, low, high, dtype=np.float32):
        self.low, self.high, self.dtype = np.asarray(low), np.asarray(high), dtype
        self.shape = self.low.shape


def to_pth_space(observation_space, action_space):
    """gym-style spaces -> the C ABI's pth_space."""
    from . import _lib
    if isinstance(action_space, Discrete):
        heads = [action_space.n]
    elif isinstance(action_space, MultiDiscrete):
        heads = action_space.nvec.tolist()
    else:
        raise ValueError(f"unsupported action space {action_space!r}")
    if isinstance(observation_space, Discrete):
        return _lib.Space.onehot([observation_space.n], heads)
    if isinstance(observation_space, MultiDiscrete):
        return _lib.Space.onehot(observation_space.nvec.tolist(), heads)
    if isinstance(observation_space, Box):
        return _lib.Space.box(int(np.prod(observation_space.shape)), heads)
    raise ValueError(f"unsupported observation space {observation_space!r}")
