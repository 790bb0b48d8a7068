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


def reference_hands(np_random):
    """Both players' dice as the reference rolls them (liar.py:22-26 `randRoll`, called for the ego and
    then for the partner, liar.py:98-99): per hand six `randint(6)` draws, returned as 2 x 6 face counts."""
    hands = []
    for _ in range(2):
        dice = [np_random.randint(N_SIDES) for _ in range(N_DICE)]
        hands += [dice.count(f) for f in range(N_SIDES)]
    return np.array(hands, np.uint8)


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

    def __init__(self, probegostart=0.5, seed=0, device="cuda", rng=None):
        """rng: "philox" (device streams) or "reference" (np.random, in the reference's draw order);
        default: pantheonrl_b200.rng_mode.get_rng_mode()."""
        super().__init__(probegostart=probegostart)
        from ..rng_mode import get_rng_mode
        self.observation_space, self.action_space = OBS_SPACE, ACTION_SPACE
        self.device, self.seed = device, int(seed)
        self.rng = rng or get_rng_mode()
        self.episodes = 0
        self.state = None  # device record, created by the first reset

    # who starts: the coin of the device reset at this episode's counter
    def draw_ego_first(self):
        if self.rng == "reference":
            return np.random.rand() < self.probegostart  # multiagentenv.py:325
        _, ego_first, _ = ops.liar_reset(1, self.seed, self.episodes, probegostart=self.probegostart,
                                         device=self.device)
        return bool(ego_first.item())

    def multi_reset(self, egofirst):
        if self.rng == "reference":
            # liar.py:22-26, 97-100: six np.random.randint(6) per hand, the ego's hand first
            hands = reference_hands(np.random)
            st = np.zeros((1, 32), np.uint8)
            st[0, :12] = hands
            self.state = torch.from_numpy(st).to(self.device)
            self.episodes += 1
            return np.concatenate([hands[:6] if egofirst else hands[6:], np.tile([N_SIDES, 0], MAX_MOVES)]
                                  ).astype(np.int64)
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
