"""Liar's Dice as a TurnBasedEnv (API of pantheonrl/envs/liargym/liar.py).  State
lives in a device record; every rule evaluation is pth_env_liar_reset /
pth_env_liar_step with N = 1 — the same device functions the rollout megakernel
inlines.  Dice and the who-starts coin come from the Philox ENV stream."""
import numpy as np
import torch

from .. import ops
from ..common.agents import Agent
from ..common.multiagentenv import TurnBasedEnv
from ..spaces import MultiDiscrete

N_SIDES, N_DICE = 6, 6
MAX_MOVES = 2 * N_DICE
BLUFF = [N_SIDES, 2 * N_DICE - 1]
ACTION_SPACE = MultiDiscrete([N_SIDES + 1, 2 * N_DICE])
OBS_SPACE = MultiDiscrete([N_DICE + 1] * N_SIDES + [N_SIDES + 1, 2 * N_DICE] * MAX_MOVES)


class LiarDefaultAgent(Agent):
    """Scripted partner: bids its most frequent face, calls when the bid exceeds it."""

    def get_action(self, obs, record=True):
        o = np.asarray(obs.obs).tolist()
        hand = o[:N_SIDES]
        best = max(hand)
        if o[N_SIDES] != N_SIDES and o[N_SIDES + 1] > best:
            return np.array(BLUFF)
        return np.array([hand.index(best), best])

    def update(self, reward, done):
        pass


class LiarEnv(TurnBasedEnv):
    device_kind = "liar"

    def __init__(self, probegostart=0.5, seed=0, device="cuda"):
        super().__init__(probegostart=probegostart)
        self.observation_space, self.action_space = OBS_SPACE, ACTION_SPACE
        self.device, self.seed = device, int(seed)
        self.episodes = 0
        self.state = torch.zeros(1, 32, dtype=torch.uint8, device=device)

    # who starts: the coin of the device reset at this episode's counter
    def draw_ego_first(self):
        _, ego_first, _ = ops.liar_reset(1, self.seed, self.episodes, probegostart=self.probegostart,
                                         device=self.device)
        return bool(ego_first.item())

    def multi_reset(self, egofirst):
        # same (seed, episode) counter -> same dice; the coin is forced to the caller's choice
        state, _, obs = ops.liar_reset(1, self.seed, self.episodes, probegostart=1.0 if egofirst else 0.0,
                                       device=self.device)
        self.state = state
        self.episodes += 1
        return obs[0, :30].cpu().numpy().astype(np.int64)

    def _step(self, action, is_ego):
        a = torch.tensor(np.asarray(action, dtype=np.uint8).reshape(1, 2), device=self.device)
        who = torch.tensor([1 if is_ego else 0], dtype=torch.uint8, device=self.device)
        obs, re, ra, done = ops.liar_step(self.state, who, a)
        return (obs[0, :30].cpu().numpy().astype(np.int64), (float(re.item()), float(ra.item())),
                bool(done.item()), {})

    def ego_step(self, action):
        return self._step(action, True)

    def alt_step(self, action):
        return self._step(action, False)

    @property
    def hands(self):
        s = self.state[0].cpu().numpy()
        return s[:6].tolist(), s[6:12].tolist()
