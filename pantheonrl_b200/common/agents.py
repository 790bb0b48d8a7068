"""Partner-agent plugin API (pantheonrl/common/agents.py): Agent, StaticPolicyAgent,
OnPolicyAgent.  Same constructor signatures and get_action / update semantics;
the arithmetic (forward + sample, GAE, PPO.train) runs in libpantheon_b200.so.
"""
from abc import ABC, abstractmethod
from collections import deque


class Agent(ABC):
    """Base class of everything MultiAgentEnv can call (agents.py:24-51)."""

    @abstractmethod
    def get_action(self, obs, record=True):
        """Return an action for this Observation; record=True while training."""

    @abstractmethod
    def update(self, reward, done):
        """Reward / done for the most recent recorded action; repeated calls add up
        and the last done flag wins."""


class StaticPolicyAgent(Agent):
    """A frozen policy (agents.py:54-79): FIXED partners and self-play."""

    def __init__(self, policy):
        self.policy = policy

    def get_action(self, obs, record=True):
        actions, _, _ = self.policy.forward(obs.obs)
        return actions[0]

    def update(self, reward, done):
        pass


class OnPolicyAgent(Agent):
    """A PPO learner playing the partner's seat (agents.py:82-208).

    It mirrors SB3's collect_rollouts one decision at a time: every recorded
    get_action appends a row to the model's rollout buffer, update() adds the
    reward to the newest row and latches `done` as the next episode_start, and
    when the buffer holds n_steps rows the NEXT get_action first runs GAE — with
    the value of the last stored observation as bootstrap (agents.py:127-129) —
    and model.train().
    """

    def __init__(self, model, log_interval=None, tensorboard_log=None, tb_log_name="OnPolicyAgent"):
        self.model = model
        self._last_episode_starts = [True]
        self.n_steps = 0
        self.values = None
        # agents.py:102-103: the partner logs through its own SB3-style logger / run directory
        from ..logger import configure_logger
        self.model.set_logger(configure_logger(getattr(model, "verbose", 0), tensorboard_log, tb_log_name))
        self.name = tb_log_name
        self.num_timesteps = 0
        self.log_interval = log_interval or (1 if getattr(model, "verbose", 0) else None)
        self.iteration = 0
        self.model.ep_info_buffer = deque([{"r": 0, "l": 0}], maxlen=100)

    def get_action(self, obs, record=True):
        buf = self.model.rollout_buffer
        if record and self.n_steps >= self.model.n_steps:
            buf.compute_returns_and_advantage(last_values=self.values, dones=self._last_episode_starts[0])
            if self.log_interval is not None and self.iteration % self.log_interval == 0:
                self._log()
            self.model.train()
            self.iteration += 1
            buf.reset()
            self.n_steps = 0
        actions, values, log_probs = self.model.policy.forward(obs.obs)
        if record:
            self.model.ep_info_buffer[-1]["l"] += 1
            buf.add(self._stored_obs(obs.obs), actions, 0.0, self._last_episode_starts[0], values, log_probs)
        self.n_steps += 1
        self.num_timesteps += 1
        self.values = values
        return actions[0]

    def _stored_obs(self, obs):
        """The row the buffer keeps for this decision (AdapAgent appends the policy's context)."""
        return obs

    def update(self, reward, done):
        self._last_episode_starts = [done]
        self.model.rollout_buffer.add_reward(reward)
        self.model.ep_info_buffer[-1]["r"] += reward
        if done:
            self.model.ep_info_buffer.append({"r": 0, "l": 0})

    def _log(self):
        """agents.py:132-153: keys, exclusions and the dump step are the reference's."""
        from ..logger import safe_mean
        lg = self.model.logger
        lg.record("name", self.name, exclude="tensorboard")
        lg.record("time/iterations", self.iteration, exclude="tensorboard")
        buf = self.model.ep_info_buffer
        if len(buf) > 0 and len(buf[0]) > 0:
            last_exclude = buf.pop()  # the episode still in progress
            lg.record("rollout/ep_rew_mean", safe_mean(ep["r"] for ep in buf))
            lg.record("rollout/ep_len_mean", safe_mean(ep["l"] for ep in buf))
            buf.append(last_exclude)
        lg.record("time/total_timesteps", self.num_timesteps, exclude="tensorboard")
        lg.dump(step=self.num_timesteps)

    def learn(self, **kwargs):
        self.model._custom_logger = False  # agents.py:206: let learn() configure its own logger
        self.model.learn(**kwargs)
