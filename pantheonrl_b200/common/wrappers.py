"""Frame-stack and recorder wrappers around the two-player env classes.

Same names, constructor arguments and recorded data as pantheonrl/common/wrappers.py:
``frame_wrap`` / ``recorder_wrap`` (:23-34), ``HistoryQueue`` (:37-74),
``TurnBasedRecorder`` (:84-165), ``SimultaneousRecorder`` (:168-232),
``TurnBasedFrameStack`` (:235-303), ``SimultaneousFrameStack`` (:306-349).
They are host-side glue for the n_envs = 1 flow (trainer.py:92-100, 364-370 enables them
with --framestack / --record); the arithmetic of the wrapped env and of the agents still runs
in libpantheon_b200.so.
"""
import numpy as np

from ..spaces import Box, Discrete, MultiDiscrete
from .multiagentenv import SimultaneousEnv, TurnBasedEnv
from .trajsaver import SimultaneousTransitions, TurnBasedTransitions

# TurnBasedRecorder flags
EGO_NOT_DONE, ALT_NOT_DONE, EGO_DONE, ALT_DONE = 0, 1, 2, 3
# SimultaneousRecorder flags
NOT_DONE, DONE = 0, 1


def calculate_space(space, numframes):
    """Observation space of ``numframes`` stacked observations (util.py:32-45)."""
    if isinstance(space, Box):
        return Box(np.tile(space.low, numframes), np.tile(space.high, numframes), dtype=space.dtype)
    if isinstance(space, Discrete):
        return MultiDiscrete([space.n] * numframes)
    if isinstance(space, MultiDiscrete):
        return MultiDiscrete(list(space.nvec) * numframes)
    raise ValueError(f"cannot stack observations of {space!r}")


def get_default_obs(env):
    """What an empty frame looks like (util.py:48-60)."""
    space = env.observation_space
    if isinstance(space, Box):
        return space.low
    if isinstance(space, Discrete):
        return [0]
    if isinstance(space, MultiDiscrete):
        return [0] * len(space.nvec)
    raise ValueError(f"no default observation for {space!r}")


def frame_wrap(env, numframes):
    return (TurnBasedFrameStack if isinstance(env, TurnBasedEnv) else SimultaneousFrameStack)(env, numframes)


def recorder_wrap(env):
    return (TurnBasedRecorder if isinstance(env, TurnBasedEnv) else SimultaneousRecorder)(env)


class HistoryQueue:
    """The last ``size`` observations, newest first, as one flat array; empty places hold
    ``defaultelem``."""

    def __init__(self, defaultelem, size):
        self.defaultelem, self.size = defaultelem, size
        self.reset()

    def reset(self):
        self.history, self.pos = [self.defaultelem] * self.size, 0

    def add(self, toadd):
        self.history[self.pos] = toadd
        newest_first = [self.history[self.pos - back] for back in range(self.size)]
        self.pos = (self.pos + 1) % self.size
        return np.array([x for frame in newest_first for x in frame])


class MultiRecorder:
    def get_transitions(self):
        raise NotImplementedError


def _adopt(wrapper, env):
    wrapper.env = env
    wrapper.action_space = env.action_space
    wrapper.observation_space = env.observation_space
    if isinstance(env, TurnBasedEnv):
        # the who-starts coin stays the wrapped game's (the reference draws it from the global
        # np.random at the outermost env with the same probability, multiagentenv.py:325; the
        # built-in games here draw it from their own reproducible stream)
        wrapper.draw_ego_first = env.draw_ego_first


class TurnBasedRecorder(TurnBasedEnv, MultiRecorder):
    """Records every observation a player acted on and the action it took.  A reset
    observation nobody has acted on yet (the run stopped, or the env was reset twice) is
    overwritten / dropped."""

    def __init__(self, env):
        super().__init__(probegostart=env.probegostart, partners=env.partners[0])
        _adopt(self, env)
        self.allobs, self.allacts, self.flags, self.incomplete = [], [], [], False

    def _moved(self, out, action, going, ended):
        nextobs, _rews, done, _info = out
        self.allacts.append(action)
        if done:
            self.flags.append(ended)
            self.incomplete = False
        else:
            self.allobs.append(nextobs)
            self.flags.append(going)
        return out

    def ego_step(self, action):
        return self._moved(self.env.ego_step(action), action, EGO_NOT_DONE, EGO_DONE)

    def alt_step(self, action):
        return self._moved(self.env.alt_step(action), action, ALT_NOT_DONE, ALT_DONE)

    def multi_reset(self, egofirst):
        first = self.env.multi_reset(egofirst)
        if self.incomplete:
            self.allobs[-1] = first
        else:
            self.allobs.append(first)
        self.incomplete = True
        return first

    def get_transitions(self):
        obs = np.array(self.allobs)
        return TurnBasedTransitions(obs[:-1] if self.incomplete else obs, np.array(self.allacts), np.array(self.flags))


class SimultaneousRecorder(SimultaneousEnv, MultiRecorder):
    """Records both players' observation / action per joint step."""

    def __init__(self, env):
        super().__init__(partners=env.partners[0])
        _adopt(self, env)
        self.allegoobs, self.allegoacts, self.allaltobs, self.allaltacts = [], [], [], []
        self.allflags, self.incomplete = [], False

    def multi_step(self, ego_action, alt_action):
        out = self.env.multi_step(ego_action, alt_action)
        obs, _rews, done, _info = out
        self.allegoacts.append(ego_action)
        self.allaltacts.append(alt_action)
        if done:
            self.allflags.append(DONE)
            self.incomplete = False
        else:
            self.allegoobs.append(obs[0])
            self.allaltobs.append(obs[1])
            self.allflags.append(NOT_DONE)
        return out

    def multi_reset(self):
        obs = self.env.multi_reset()
        self.allegoobs.append(obs[0])
        self.allaltobs.append(obs[1])
        self.incomplete = True
        return obs

    def get_transitions(self):
        ego, alt = np.array(self.allegoobs), np.array(self.allaltobs)
        if self.incomplete:
            ego, alt = ego[:-1], alt[:-1]
        return SimultaneousTransitions(ego, np.array(self.allegoacts), alt, np.array(self.allaltacts),
                                       np.array(self.allflags))


class TurnBasedFrameStack(TurnBasedEnv):
    """Each player sees its last ``numframes`` observations, newest first."""

    def __init__(self, env, numframes, defaultobs=None, altenv=None, defaultaltobs=None):
        super().__init__(probegostart=env.probegostart, partners=env.partners[0])
        _adopt(self, env)
        self.numframes = numframes
        self.observation_space = calculate_space(env.observation_space, numframes)
        ego_default = defaultobs if defaultobs is not None else get_default_obs(env)
        alt_default = defaultaltobs if defaultaltobs is not None else get_default_obs(altenv if altenv is not None else env)
        self.egohistory = HistoryQueue(ego_default, numframes)
        self.althistory = HistoryQueue(alt_default, numframes)

    def ego_step(self, action):
        altobs, rews, done, info = self.env.ego_step(action)
        return self.althistory.add(altobs), rews, done, info

    def alt_step(self, action):
        egoobs, rews, done, info = self.env.alt_step(action)
        return self.egohistory.add(egoobs), rews, done, info

    def multi_reset(self, egofirst):
        first = self.env.multi_reset(egofirst)
        self.egohistory.reset()
        self.althistory.reset()
        return (self.egohistory if egofirst else self.althistory).add(first)


class SimultaneousFrameStack(SimultaneousEnv):
    """Both players see their last ``numframes`` observations, newest first."""

    def __init__(self, env, numframes, defaultobs=None):
        super().__init__(partners=env.partners[0])
        _adopt(self, env)
        self.numframes = numframes
        self.observation_space = calculate_space(env.observation_space, numframes)
        self.defaultobs = get_default_obs(env) if defaultobs is None else list(defaultobs)
        self.egohistory = HistoryQueue(self.defaultobs, numframes)
        self.althistory = HistoryQueue(self.defaultobs, numframes)

    def multi_step(self, ego_action, alt_action):
        obs, rews, done, info = self.env.multi_step(ego_action, alt_action)
        return (self.egohistory.add(obs[0]), self.althistory.add(obs[1])), rews, done, info

    def multi_reset(self):
        obs = self.env.multi_reset()
        self.egohistory.reset()
        self.althistory.reset()
        return self.egohistory.add(obs[0]), self.althistory.add(obs[1])
