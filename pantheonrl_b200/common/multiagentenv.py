"""Ego-centric multi-agent env shim with the API of
pantheonrl/common/multiagentenv.py (MultiAgentEnv :25-284, TurnBasedEnv :287-380,
SimultaneousEnv :383-442).  Written from the behaviour the reference defines:
who moves when, partner selection per episode, reward routing with the
first-move hand-off, and returning the previous ego observation on the terminal
step.  tests/test_gpu_facade.py replays traces recorded from the reference's own
classes through this file.

This is the N = 1, host-driven way to run the path (one Python call per
decision, each agent call hitting the CUDA kernels with B = 1).  The N >> 1 way
is pantheonrl_b200.engine.VecTrainer, where the same logic runs inside
pth_rollout_run.
"""
from abc import ABC, abstractmethod

import numpy as np

from .observation import Observation, extract_obs


class PlayerException(Exception):
    """Players of the environment are set up incorrectly."""


class DummyEnv:
    """Spaces-only stand-in a partner can build its policy from."""

    def __init__(self, observation_space, action_space):
        self.observation_space, self.action_space = observation_space, action_space


class MultiAgentEnv(ABC):
    def __init__(self, ego_ind=0, n_players=2, resample_policy="default", partners=None,
                 ego_extractor=extract_obs):
        self.ego_ind, self.n_players = ego_ind, n_players
        if partners is not None:
            if len(partners) != n_players - 1:
                raise PlayerException("The number of partners needs to equal the number of non-ego players")
            if any((not isinstance(pl, list)) or len(pl) == 0 for pl in partners):
                raise PlayerException("Sublist for each partner must be nonempty list")
        # the reference aliases ONE list for every slot when none are given (multiagentenv.py:57)
        self.partners = partners if partners else [[]] * (n_players - 1)
        self.partnerids = [0] * (n_players - 1)
        self._players, self._obs, self._old_ego_obs = (), (), None
        self.should_update = [False] * (n_players - 1)
        self.total_rews = [0] * n_players
        self.ego_moved = False
        self.set_resample_policy(resample_policy)
        self.ego_extractor = ego_extractor

    # ------------------------------------------------------------ partners
    def getDummyEnv(self, player_num):
        return self

    def set_ego_extractor(self, ego_extractor):
        self.ego_extractor = ego_extractor

    def _get_partner_num(self, player_num):
        if player_num == self.ego_ind:
            raise PlayerException("Ego agent is not set by the environment")
        return player_num - 1 if player_num > self.ego_ind else player_num

    def add_partner_agent(self, agent, player_num=1):
        self.partners[self._get_partner_num(player_num)].append(agent)

    def set_partnerid(self, agent_id, player_num=1):
        slot = self._get_partner_num(player_num)
        assert 0 <= agent_id < len(self.partners[slot])
        self.partnerids[slot] = agent_id

    def resample_random(self):
        self.partnerids = [np.random.randint(len(pl)) for pl in self.partners]

    def resample_round_robin(self):
        self.partnerids = [(self.partnerids[0] + 1) % len(self.partners[0])]

    def set_resample_policy(self, resample_policy):
        if resample_policy == "default":
            resample_policy = "robin" if self.n_players == 2 else "random"
        if resample_policy == "robin":
            if self.n_players != 2:
                raise PlayerException("Cannot do round robin resampling for >2 players")
            self.resample_partner = self.resample_round_robin
        elif resample_policy == "random":
            self.resample_partner = self.resample_random
        else:
            raise PlayerException(f"Invalid resampling policy: {resample_policy}")

    # ------------------------------------------------------------ routing
    def _current_partner(self, player):
        slot = self._get_partner_num(player)
        return slot, self.partners[slot][self.partnerids[slot]]

    def _get_actions(self, players, obs, ego_act=None):
        acts = []
        for player, ob in zip(players, obs):
            if player == self.ego_ind:
                acts.append(ego_act)
                continue
            slot, agent = self._current_partner(player)
            acts.append(agent.get_action(ob))
            if not self.should_update[slot]:
                # first move of the episode: hand over what accrued before it moved
                agent.update(self.total_rews[player], False)
                self.should_update[slot] = True
        return np.array(acts)

    def _update_players(self, rews, done):
        for slot in range(self.n_players - 1):
            player = slot if slot < self.ego_ind else slot + 1
            if self.should_update[slot]:
                self.partners[slot][self.partnerids[slot]].update(rews[player], done)
        for i in range(self.n_players):
            self.total_rews[i] += rews[i]

    def step(self, action):
        """One ego timestep: play until the ego is to move again or the game ends."""
        ego_rew = 0.0
        while True:
            acts = self._get_actions(self._players, self._obs, action)
            self._players, self._obs, rews, done, info = self.n_step(acts)
            info["_partnerid"] = self.partnerids
            self._update_players(rews, done)
            ego_rew += rews[self.ego_ind] if self.ego_moved else self.total_rews[self.ego_ind]
            self.ego_moved = True
            if done:
                return self.ego_extractor(self._old_ego_obs), ego_rew, done, info
            if self.ego_ind in self._players:
                break
        ego_obs = self._obs[self._players.index(self.ego_ind)]
        self._old_ego_obs = ego_obs
        return self.ego_extractor(ego_obs), ego_rew, done, info

    def reset(self):
        self.resample_partner()
        self._players, self._obs = self.n_reset()
        self.should_update = [False] * (self.n_players - 1)
        self.total_rews = [0] * self.n_players
        self.ego_moved = False
        while self.ego_ind not in self._players:
            acts = self._get_actions(self._players, self._obs)
            self._players, self._obs, rews, done, _ = self.n_step(acts)
            if done:
                raise PlayerException("Game ended before ego moved")
            self._update_players(rews, done)
        ego_obs = self._obs[self._players.index(self.ego_ind)]
        assert ego_obs is not None
        self._old_ego_obs = ego_obs
        return self.ego_extractor(ego_obs)

    # gym.Env's housekeeping surface, as callers of the reference use it (tester.py:47-58)
    def close(self):
        pass

    def render(self, mode="human"):
        pass

    def seed(self, seed=None):
        return [seed]

    @abstractmethod
    def n_step(self, actions):
        """-> (next players, their observations, rewards of all players, done, info)"""

    @abstractmethod
    def n_reset(self):
        """-> (players that move first, their observations)"""


class TurnBasedEnv(MultiAgentEnv, ABC):
    """Two players alternate; subclasses give ego_step / alt_step / multi_reset."""

    def __init__(self, probegostart=0.5, partners=None):
        super().__init__(ego_ind=0, n_players=2, partners=[partners] if partners else None)
        self.probegostart = probegostart
        self.ego_next = True

    def n_step(self, actions):
        mover_is_ego = self.ego_next
        obs, rews, done, info = (self.ego_step if mover_is_ego else self.alt_step)(actions[0])
        self.ego_next = not mover_is_ego
        return (1 if mover_is_ego else 0,), (Observation(obs),), rews, done, info

    def n_reset(self):
        self.ego_next = bool(self.draw_ego_first())
        obs = self.multi_reset(self.ego_next)
        return (0 if self.ego_next else 1,), (Observation(obs),)

    def draw_ego_first(self):
        """np.random.rand() < probegostart in the reference (multiagentenv.py:325)."""
        return np.random.rand() < self.probegostart

    @abstractmethod
    def ego_step(self, action):
        """-> (partner's obs, (ego reward, partner reward), done, info)"""

    @abstractmethod
    def alt_step(self, action):
        """-> (ego's obs, (ego reward, partner reward), done, info)"""

    @abstractmethod
    def multi_reset(self, egofirst):
        """-> observation of the player that starts"""


class SimultaneousEnv(MultiAgentEnv, ABC):
    """Both players act on every step; subclasses give multi_step / multi_reset."""

    def __init__(self, partners=None):
        super().__init__(ego_ind=0, n_players=2, partners=[partners] if partners else None)

    def n_step(self, actions):
        (o0, o1), rews, done, info = self.multi_step(actions[0], actions[1])
        return (0, 1), (Observation(o0), Observation(o1)), rews, done, info

    def n_reset(self):
        o0, o1 = self.multi_reset()
        return (0, 1), (Observation(o0), Observation(o1))

    @abstractmethod
    def multi_step(self, ego_action, alt_action):
        """-> ((ego obs, partner obs), (ego reward, partner reward), done, info)"""

    @abstractmethod
    def multi_reset(self):
        """-> (ego obs, partner obs)"""
