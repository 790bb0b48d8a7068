"""Rock-Paper-Scissors as a SimultaneousEnv (API of pantheonrl/envs/rpsgym/rps.py).
The payoff is computed by pth_env_rps_step (N = 1), the same kernel code the
rollout megakernel inlines."""
import numpy as np
import torch

from .. import ops
from ..common.agents import Agent
from ..common.multiagentenv import SimultaneousEnv
from ..spaces import Discrete

ACTION_NAMES = ["ROCK", "PAPER", "SCISSORS"]
ACTION_SPACE = Discrete(3)
OBS_SPACE = Discrete(1)
NULL_OBS = np.array([0])


class RPSWeightedAgent(Agent):
    """Scripted partner: plays rock/paper/scissors with fixed weights."""

    def __init__(self, r=1, p=1, s=1, np_random=np.random):
        w = r + p + s
        self.c0, self.c1 = (1 / 3, 2 / 3) if w == 0 else (r / w, (r + p) / w)
        self.np_random = np_random

    def get_action(self, obs, record=True):
        roll = self.np_random.rand()
        return 0 if roll < self.c0 else (1 if roll < self.c1 else 2)

    def update(self, reward, done):
        pass


class RPSEnv(SimultaneousEnv):
    device_kind = "rps"  # on-device twin for n_envs > 1 (pth_rollout_run)

    def __init__(self, device="cuda"):
        super().__init__()
        self.history = []
        self.observation_space, self.action_space = OBS_SPACE, ACTION_SPACE
        self.device = device

    def multi_step(self, ego_action, alt_action):
        a = torch.tensor([int(ego_action)], dtype=torch.int32, device=self.device)
        b = torch.tensor([int(alt_action)], dtype=torch.int32, device=self.device)
        re, ra = ops.rps_step(a, b)
        return (NULL_OBS, NULL_OBS), (int(re.item()), int(ra.item())), True, {}

    def multi_reset(self):
        return NULL_OBS, NULL_OBS
