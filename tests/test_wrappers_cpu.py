"""Frame-stack / recorder wrappers and the .npy trajectory format against
fixtures produced by the reference's OWN pantheonrl/common/wrappers.py + trajsaver.py
(tests/golden/make_golden_wrappers.py): same scripts, same dice, and every observation the
ego and the partner see, every recorded row and the written .npy file must be identical."""
import io
import os

import numpy as np
import pytest

import oracle
from pantheonrl_b200.common import trajsaver, wrappers
from pantheonrl_b200.common.agents import Agent
from pantheonrl_b200.common.multiagentenv import SimultaneousEnv, TurnBasedEnv
from pantheonrl_b200.spaces import Box, Discrete, MultiDiscrete


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "wrappers.npz"))


class Script(Agent):
    def __init__(self, actions):
        self.actions, self.k, self.seen = actions, 0, []

    def get_action(self, obs, record=True):
        self.seen.append(np.asarray(obs.obs).reshape(-1).copy())
        a = self.actions[self.k % len(self.actions)]
        self.k += 1
        return a

    def update(self, reward, done):
        pass


class ReplayLiar(TurnBasedEnv):
    """Liar's Dice on the CPU oracle's rules (pinned on the reference's LiarEnv elsewhere) with
    the coin and the dice of every reset taken from the fixture."""

    def __init__(self, resets):
        super().__init__()
        self.observation_space = MultiDiscrete([7] * 6 + [7, 12] * 12)
        self.action_space = MultiDiscrete([7, 12])
        self.resets, self.k = resets, 0
        self.state = np.zeros((1, 32), np.uint8)

    def draw_ego_first(self):
        return bool(self.resets[self.k][0])

    def multi_reset(self, egofirst):
        row = self.resets[self.k]
        self.k += 1
        assert bool(row[0]) == bool(egofirst)
        self.state[:] = 0
        self.state[0, :12] = row[1:13]
        hand = row[1:7] if egofirst else row[7:13]
        return np.array(list(hand) + [6, 0] * 12)

    def _move(self, action, is_ego):
        obs, re, ra, done = oracle.liar_step(self.state, [int(is_ego)], np.asarray(action, np.uint8).reshape(1, 2))
        return obs[0, :30].astype(np.int64), (float(re[0]), float(ra[0])), bool(done[0]), {}

    def ego_step(self, action):
        return self._move(action, True)

    def alt_step(self, action):
        return self._move(action, False)


class CountingGame(SimultaneousEnv):
    def __init__(self, length=3):
        super().__init__()
        self.length = length
        self.observation_space = MultiDiscrete([length + 1, 4, 4])
        self.action_space = Discrete(3)

    def multi_reset(self):
        self.t = 0
        o = np.array([0, 3, 3])
        return o, o

    def multi_step(self, ego_action, alt_action):
        self.t += 1
        o = (int(ego_action) - int(alt_action) + 3) % 3
        r = -1 if o == 2 else o
        return (np.array([self.t, int(ego_action), int(alt_action)]),
                np.array([self.t, int(alt_action), int(ego_action)])), (r, -r), self.t >= self.length, {}


def run(env, ego_actions, n_steps):
    obs_seen, rews, dones = [], [], []
    o = env.reset()
    obs_seen.append(np.asarray(o).reshape(-1).copy())
    for t in range(n_steps):
        o, r, d, _ = env.step(ego_actions[t % len(ego_actions)])
        rews.append(r)
        dones.append(d)
        if d:
            o = env.reset()
        obs_seen.append(np.asarray(o).reshape(-1).copy())
    return np.array(obs_seen), np.array(rews, np.float64), np.array(dones)


def npy_bytes(tr):
    b = io.BytesIO()
    tr.write_transition(b)
    return np.frombuffer(b.getvalue(), np.uint8)


@pytest.mark.parametrize("name,order", [("liar_fr", ("frame", "record")), ("liar_rf", ("record", "frame"))])
def test_turn_based_wrappers_match_the_reference(g, name, order):
    env = ReplayLiar(g[f"{name}_resets"])
    partner = Script(list(g["liar_alt_script"]))
    env.add_partner_agent(partner)
    rec = None
    for w in order:
        if w == "frame":
            env = wrappers.frame_wrap(env, 3)
            assert isinstance(env, wrappers.TurnBasedFrameStack)
        else:
            env = rec = wrappers.recorder_wrap(env)
            assert isinstance(rec, wrappers.TurnBasedRecorder)
    assert np.array_equal(env.observation_space.nvec, g[f"{name}_obs_nvec"])
    obs, rews, dones = run(env, list(g["liar_ego_script"]), 60)
    assert np.array_equal(obs, g[f"{name}_ego_obs"])          # what the ego saw (stacked, newest first)
    assert np.array_equal(rews, g[f"{name}_rews"]) and np.array_equal(dones, g[f"{name}_dones"])
    assert np.array_equal(np.array(partner.seen), g[f"{name}_partner_seen"])
    tr = rec.get_transitions()
    assert np.array_equal(tr.obs, g[f"{name}_rec_obs"]) and np.array_equal(tr.acts, g[f"{name}_rec_acts"])
    assert np.array_equal(tr.flags, g[f"{name}_rec_flags"])
    assert len(tr.get_ego_transitions()) == int(g[f"{name}_ego_n"])
    assert len(tr.get_alt_transitions()) == int(g[f"{name}_alt_n"])
    assert np.array_equal(npy_bytes(tr), g[f"{name}_npy"])    # the file itself, byte for byte
    back = trajsaver.TurnBasedTransitions.read_transition(io.BytesIO(g[f"{name}_npy"].tobytes()),
                                                          rec.observation_space, rec.action_space)
    assert np.array_equal(back.obs, tr.obs) and np.array_equal(back.acts, tr.acts) and np.array_equal(back.flags, tr.flags)


def test_simultaneous_wrappers_match_the_reference(g):
    env = CountingGame(3)
    partner = Script(list(g["sim_alt_script"]))
    env.add_partner_agent(partner)
    env = wrappers.frame_wrap(env, 2)
    assert isinstance(env, wrappers.SimultaneousFrameStack)
    env = rec = wrappers.recorder_wrap(env)
    assert isinstance(rec, wrappers.SimultaneousRecorder)
    assert np.array_equal(env.observation_space.nvec, g["sim_obs_nvec"])
    obs, rews, dones = run(env, list(g["sim_ego_script"]), 20)
    assert np.array_equal(obs, g["sim_ego_obs"]) and np.array_equal(rews, g["sim_rews"])
    assert np.array_equal(dones, g["sim_dones"]) and np.array_equal(np.array(partner.seen), g["sim_partner_seen"])
    tr = rec.get_transitions()
    for ours, ref in ((tr.egoobs, "sim_egoobs"), (tr.egoacts, "sim_egoacts"), (tr.altobs, "sim_altobs"),
                      (tr.altacts, "sim_altacts"), (tr.flags, "sim_flags")):
        assert np.array_equal(ours, g[ref]), ref
    assert np.array_equal(npy_bytes(tr), g["sim_npy"])
    back = trajsaver.SimultaneousTransitions.read_transition(io.BytesIO(g["sim_npy"].tobytes()),
                                                             env.observation_space, env.action_space)
    assert np.array_equal(back.altobs, tr.altobs) and np.array_equal(back.egoacts.reshape(-1), tr.egoacts)
    ego = tr.get_ego_transitions()
    assert ego[3]["obs"].shape == (6,) and len(ego[2:5]) == 3


def test_history_queue_spaces_and_minimal_transitions(tmp_path):
    q = wrappers.HistoryQueue([0, 0], 3)
    assert q.add([1, 2]).tolist() == [1, 2, 0, 0, 0, 0]
    assert q.add([3, 4]).tolist() == [3, 4, 1, 2, 0, 0]
    q.add([5, 6])
    assert q.add([7, 8]).tolist() == [7, 8, 5, 6, 3, 4]      # the oldest frame fell out
    q.reset()
    assert q.add([9, 9]).tolist() == [9, 9, 0, 0, 0, 0]
    assert wrappers.calculate_space(Discrete(3), 4).nvec.tolist() == [3, 3, 3, 3]
    box = wrappers.calculate_space(Box(np.zeros(2), np.ones(2)), 2)
    assert box.shape == (4,) and box.high.tolist() == [1, 1, 1, 1]
    assert wrappers.get_default_obs(type("E", (), {"observation_space": Discrete(5)})()) == [0]
    assert trajsaver.get_space_size(MultiDiscrete([7, 12])) == 2 and trajsaver.get_space_size(Discrete(3)) == 1
    tm = trajsaver.TransitionsMinimal(np.arange(12).reshape(4, 3), np.arange(4).reshape(4, 1))
    tm.write_transition(tmp_path / "t.npy")
    back = trajsaver.TransitionsMinimal.read_transition(tmp_path / "t.npy", MultiDiscrete([9, 9, 9]), Discrete(4))
    assert np.array_equal(back.obs, tm.obs) and np.array_equal(back.acts, tm.acts)
    with pytest.raises(ValueError):
        trajsaver.TransitionsMinimal(np.zeros((3, 2)), np.zeros((2, 1)))
    with pytest.raises(ValueError):
        tm.obs[0, 0] = 1  # arrays are made read-only, like the reference's
