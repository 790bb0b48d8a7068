"""SB3-shaped checkpoint archives: layout, round trip, and reading an
archive rendered the way a real SB3 1.7.0 renders it (pickled entries never unpickled)."""
import base64
import io
import json
import pickle
import zipfile

import numpy as np
import pytest
import torch

from pantheonrl_b200 import _lib, checkpoint as ck, policy as pol
from pantheonrl_b200.spaces import Box, Discrete, MultiDiscrete, to_pth_space

LIAR_OBS = MultiDiscrete([7] * 6 + [7, 12] * 12)
LIAR_ACT = MultiDiscrete([7, 12])


def _state(space, seed):
    flat = pol.init_flat(space, seed)
    return flat, pol.flat_to_state_dict(space, flat)


def test_layout_and_roundtrip(tmp_path):
    space = to_pth_space(LIAR_OBS, LIAR_ACT)
    flat, sd = _state(space, 10)
    names = [n for n, _ in pol.tensor_shapes(space)]
    rng = np.random.default_rng(0)
    m = pol.flat_to_state_dict(space, rng.standard_normal(flat.size).astype(np.float32))
    v = pol.flat_to_state_dict(space, rng.random(flat.size).astype(np.float32))
    hyper = dict(learning_rate=3e-4, n_steps=128, batch_size=64, n_epochs=10, gamma=0.99, gae_lambda=0.95,
                 clip_range=0.2, normalize_advantage=True, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, seed=10,
                 n_envs=4096, verbose=0, n_minibatches=32)
    path = ck.save_zip(tmp_path / "ego", LIAR_OBS, LIAR_ACT, hyper, sd,
                       ck.optimizer_state_dict(names, m, v, 640, 3e-4),
                       {"num_timesteps": 1 << 20, "n_updates": 20, "adam_step": 640})
    assert path.endswith("ego.zip")  # SB3's open_path appends the suffix
    with zipfile.ZipFile(path) as z:
        assert {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth",
                "_stable_baselines3_version", "system_info.txt"} <= set(z.namelist())
        assert z.read("_stable_baselines3_version").decode() == "1.7.0"
        raw_sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True)
        raw_opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), weights_only=True)
    # a plain torch module with SB3's parameter names can load policy.pth as is
    assert list(raw_sd) == names
    assert raw_sd["mlp_extractor.policy_net.0.weight"].shape == (64, 270)
    assert raw_sd["action_net.weight"].shape == (19, 64)
    # and torch.optim.Adam accepts the optimizer state (SB3: Adam over policy.parameters(), eps 1e-5)
    params = [torch.nn.Parameter(raw_sd[n].clone()) for n in names]
    opt = torch.optim.Adam(params, lr=1.0, eps=1e-5)
    opt.load_state_dict(raw_opt)
    assert opt.param_groups[0]["lr"] == pytest.approx(3e-4) and opt.param_groups[0]["eps"] == 1e-5
    assert float(opt.state[params[0]]["step"]) == 640
    assert torch.equal(opt.state[params[3]]["exp_avg"], m[names[3]])

    c = ck.load_zip(tmp_path / "ego")  # suffix-less path, like PPO.load(location)
    assert c["hyper"]["n_steps"] == 128 and c["hyper"]["n_minibatches"] == 32 and c["hyper"]["seed"] == 10
    assert c["counters"] == {"num_timesteps": 1 << 20, "n_updates": 20, "adam_step": 640}
    assert isinstance(c["observation_space"], MultiDiscrete) and c["observation_space"].nvec.tolist() == LIAR_OBS.nvec.tolist()
    assert c["action_space"].nvec.tolist() == [7, 12]
    assert np.array_equal(pol.state_dict_to_flat(space, c["policy"]), flat)  # bit-exact round trip
    mm, vv = ck.adam_moments(names, c["optimizer"])
    assert all(torch.equal(mm[n], m[n]) and torch.equal(vv[n], v[n]) for n in names)


def test_pickled_entries_resolve_to_gym_and_sb3_names():
    """The :serialized: blobs reference gym / SB3 classes by name (what a real SB3 needs
    to unpickle them); decode the opcodes without importing anything."""
    e = ck._space_entry(Discrete(3))
    blob = base64.b64decode(e[":serialized:"])
    ops = [(op.name, arg) for op, arg, _ in __import__("pickletools").genops(blob)]
    assert ("GLOBAL", "gym.spaces.discrete Discrete") in ops
    assert any(o == "BININT1" and a == 3 for o, a in ops)
    assert e[":type:"] == "<class 'gym.spaces.discrete.Discrete'>" and e["n"] == 3
    with pytest.raises(ModuleNotFoundError):
        pickle.loads(blob)  # no gym here -- and no stub left behind in sys.modules
    e = ck._space_entry(LIAR_OBS)
    ops = [arg for op, arg, _ in __import__("pickletools").genops(base64.b64decode(e[":serialized:"]))]
    assert "gym.spaces.multi_discrete MultiDiscrete" in ops
    blob = ck._pickle_by_reference("stable_baselines3.common.policies", "ActorCriticPolicy")
    assert b"stable_baselines3.common.policies" in blob and b"ActorCriticPolicy" in blob
    import sys
    assert "gym" not in sys.modules and "stable_baselines3" not in sys.modules


def _sb3_rendered_archive(path, obs_entry, act_entry, sd, opt):
    """An archive the way SB3 1.7.0's data_to_json renders it: non-JSON attributes are
    {":type:", ":serialized:", **str(__dict__)}; here the pickles are garbage on purpose."""
    junk = base64.b64encode(b"\x80\x04not-a-pickle").decode()
    data = {
        "policy_class": {":type:": "<class 'abc.ABCMeta'>", ":serialized:": junk,
                         "__module__": "stable_baselines3.common.policies"},
        "verbose": 1, "observation_space": obs_entry, "action_space": act_entry, "n_envs": 1,
        "num_timesteps": 4096, "_total_timesteps": 4096, "seed": None, "learning_rate": 0.0003,
        "lr_schedule": {":type:": "<class 'function'>", ":serialized:": junk},
        "clip_range": {":type:": "<class 'function'>", ":serialized:": junk},
        "_last_obs": {":type:": "<class 'numpy.ndarray'>", ":serialized:": junk},
        "ep_info_buffer": {":type:": "<class 'collections.deque'>", ":serialized:": junk},
        "_n_updates": 20, "n_steps": 2048, "gamma": 0.99, "gae_lambda": 0.95, "ent_coef": 0.0, "vf_coef": 0.5,
        "max_grad_norm": 0.5, "batch_size": 64, "n_epochs": 10, "normalize_advantage": True, "target_kl": None,
        "policy_kwargs": {},
    }
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps(data, indent=4))
        for name, obj in (("pytorch_variables.pth", None), ("policy.pth", sd), ("policy.optimizer.pth", opt)):
            b = io.BytesIO()
            torch.save(obj, b)
            z.writestr(name, b.getvalue())
        z.writestr("_stable_baselines3_version", "1.7.0")


def test_reads_sb3_rendering(tmp_path):
    junk = base64.b64encode(b"junk").decode()
    obs = {":type:": "<class 'gym.spaces.box.Box'>", ":serialized:": junk, "dtype": "float64", "_shape": [62],
           "low": str(np.zeros(62)), "high": str(np.full(62, 20.0)), "bounded_below": str(np.ones(62, bool)),
           "bounded_above": str(np.ones(62, bool)), "_np_random": None}
    act = {":type:": "<class 'gym.spaces.discrete.Discrete'>", ":serialized:": junk, "n": 6, "_shape": [],
           "dtype": "int64", "_np_random": None}
    space = _lib.Space.box(62, [6])
    flat, sd = _state(space, 3)
    names = [n for n, _ in pol.tensor_shapes(space)]
    # a torch.optim.Adam that really stepped, so the state has torch's own structure
    params = [torch.nn.Parameter(sd[n].clone()) for n in names]
    opt = torch.optim.Adam(params, lr=3e-4, eps=1e-5)
    for _ in range(3):
        for p in params:
            p.grad = torch.ones_like(p) * 0.01
        opt.step()
    _sb3_rendered_archive(tmp_path / "sb3.zip", obs, act, {n: p.detach() for n, p in zip(names, params)},
                          opt.state_dict())
    c = ck.load_zip(tmp_path / "sb3.zip")
    assert isinstance(c["observation_space"], Box) and c["observation_space"].shape == (62,)
    assert c["observation_space"].high.max() == 20.0 and c["observation_space"].dtype == np.float64
    assert isinstance(c["action_space"], Discrete) and c["action_space"].n == 6
    assert c["hyper"]["n_steps"] == 2048 and "clip_range" not in c["hyper"] and c["hyper"]["seed"] is None
    assert c["counters"]["adam_step"] == 3 and c["counters"]["n_updates"] == 20
    mm, vv = ck.adam_moments(names, c["optimizer"])
    assert torch.equal(mm[names[0]], opt.state[params[0]]["exp_avg"])
    assert pol.state_dict_to_flat(space, c["policy"]).shape == flat.shape


def test_rejects_foreign_zip(tmp_path):
    with zipfile.ZipFile(tmp_path / "x.zip", "w") as z:
        z.writestr("readme", "hi")
    with pytest.raises(ValueError):
        ck.load_zip(tmp_path / "x.zip")
