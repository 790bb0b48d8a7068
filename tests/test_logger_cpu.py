"""SB3-shaped logger: tags, exclusions, run directories, and the partner's
log record (pantheonrl/common/agents.py:132-153) through the facade's OnPolicyAgent."""
import io
import math
import os

import torch

from pantheonrl_b200 import logger as lg
from pantheonrl_b200.common.agents import OnPolicyAgent


class Capture:
    def __init__(self):
        self.dumps = []

    def write(self, kv, excluded, step=0):
        self.dumps.append((dict(kv), dict(excluded), step))

    def close(self):
        pass


def test_record_dump_and_exclusions():
    cap, out = Capture(), io.StringIO()
    L = lg.Logger(None, [lg.HumanOutputFormat(out), cap])
    L.record("time/iterations", 3, exclude="tensorboard")
    L.record("rollout/ep_rew_mean", 0.25)
    L.record("name", "OnPolicyAgent", exclude=("tensorboard", "stdout"))
    L.record_mean("train/x", 1.0)
    L.record_mean("train/x", 3.0)
    L.dump(step=2048)
    kv, ex, step = cap.dumps[0]
    assert step == 2048 and kv["train/x"] == 2.0 and ex["time/iterations"] == ("tensorboard",)
    text = out.getvalue()
    assert "rollout/" in text and "ep_rew_mean" in text and "| time/" in text and "OnPolicyAgent" not in text
    L.dump(step=1)
    assert cap.dumps[1][0] == {}  # cleared by dump


def test_run_directories_and_tensorboard_scalars(tmp_path):
    a = lg.configure_logger(0, str(tmp_path), "PPO")
    b = lg.configure_logger(0, str(tmp_path), "PPO")
    c = lg.configure_logger(0, str(tmp_path), "OnPolicyAgent")
    assert a.get_dir().endswith("PPO_1") and b.get_dir().endswith("PPO_2") and c.get_dir().endswith("OnPolicyAgent_1")
    assert lg.get_latest_run_id(str(tmp_path), "PPO") == 2
    a.record("train/loss", 0.5)
    a.record("time/total_timesteps", 100, exclude="tensorboard")
    a.record("name", "x")  # strings never reach TensorBoard
    a.dump(step=100)
    a.record("train/loss", 0.25)
    a.dump(step=200)
    a.close()
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(a.get_dir())
    acc.Reload()
    assert acc.Tags()["scalars"] == ["train/loss"]
    assert [(e.step, e.value) for e in acc.Scalars("train/loss")] == [(100, 0.5), (200, 0.25)]
    assert lg.configure_logger(0, None).output_formats == []
    assert len(lg.configure_logger(1, None).output_formats) == 1


def test_helpers():
    assert math.isnan(lg.safe_mean([])) and lg.safe_mean([1, 2, 3]) == 2.0
    r = torch.tensor([1.0, 2.0, 3.0, 4.0])
    assert lg.explained_variance(r, r) == 1.0
    assert abs(lg.explained_variance(torch.zeros(4), r)) < 1e-6
    assert math.isnan(lg.explained_variance(r, torch.ones(4)))


class _StubBuffer:
    def __init__(self):
        self.rewards, self.pos = [], 0

    def add(self, *a):
        self.rewards.append(0.0)
        self.pos += 1

    def add_reward(self, r):
        if self.rewards:
            self.rewards[-1] += r

    def reset(self):
        self.rewards, self.pos = [], 0

    def compute_returns_and_advantage(self, last_values, dones):
        self.bootstrap = (last_values, dones)


class _StubPolicy:
    def forward(self, obs):
        import numpy as np
        return np.array([1]), 0.5, -1.0


class _StubModel:
    """What OnPolicyAgent reaches into (agents.py:97-109, 123-184)."""
    verbose, n_steps = 0, 4

    def __init__(self):
        self.rollout_buffer, self.policy, self.trained = _StubBuffer(), _StubPolicy(), 0

    def set_logger(self, logger):
        self.logger = logger

    def train(self):
        self.trained += 1


def test_partner_log_record(tmp_path):
    """Keys, exclusions and step of the record an OnPolicyAgent dumps before each train()."""
    from pantheonrl_b200.common.observation import Observation
    m = _StubModel()
    agent = OnPolicyAgent(m, log_interval=1, tensorboard_log=str(tmp_path), tb_log_name="partner")
    assert os.path.isdir(tmp_path / "partner_1")
    cap = Capture()
    m.logger.output_formats.append(cap)
    obs = Observation(0)
    for t in range(9):  # RPS-like: every decision ends an episode with reward +-1
        agent.get_action(obs)
        agent.update(1.0 if t % 2 == 0 else -1.0, True)
    assert m.trained == 2 and len(cap.dumps) == 2
    kv, ex, step = cap.dumps[0]
    assert step == 4 and kv["name"] == "partner" and kv["time/iterations"] == 0 and kv["time/total_timesteps"] == 4
    assert kv["rollout/ep_rew_mean"] == 0.0 and kv["rollout/ep_len_mean"] == 1.0  # 4 finished episodes: +1 -1 +1 -1
    assert ex["name"] == ("tensorboard",) and ex["rollout/ep_rew_mean"] == ()
    assert cap.dumps[1][2] == 8 and cap.dumps[1][0]["time/iterations"] == 1
