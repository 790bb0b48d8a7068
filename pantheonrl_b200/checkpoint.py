"""SB3-shaped checkpoint archives.

The reference saves and loads agents only through stable-baselines3's
``BaseAlgorithm.save`` / ``PPO.load`` (trainer.py:146-151, 419-432;
overcookedgym/overcooked-flask/app.py:187-190).  SB3 1.7.0 writes a zip with

  data                        JSON of the algorithm's attributes; anything json cannot
                              express becomes {":type:", ":serialized:" (base64 cloudpickle),
                              plus the object's __dict__ rendered as strings}
  policy.pth                  torch.save(policy.state_dict())
  policy.optimizer.pth        torch.save(policy.optimizer.state_dict())
  pytorch_variables.pth       torch.save(None-or-dict)
  _stable_baselines3_version  text
  system_info.txt             text

This module writes that layout and reads it back — our own archives and archives
written by a real SB3 (their pickled entries are never unpickled: spaces are
rebuilt from the string fields SB3 stores beside the pickle, hyper-parameters from
the plain JSON values).  It is host-side file plumbing: no device code, no numbers
on the hot path.  Tensors use torch layout (weight[out][in]) and SB3's parameter
names / order, see policy.tensor_shapes().
"""
import base64
import io
import json
import pickle
import re
import sys
import types
import zipfile

import numpy as np
import torch

from .spaces import Box, Discrete, MultiDiscrete

SB3_VERSION = "1.7.0"
HYPER_KEYS = ("learning_rate", "n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "clip_range",
              "normalize_advantage", "ent_coef", "vf_coef", "max_grad_norm", "seed", "n_envs", "verbose")
_GYM_TYPES = {Discrete: "gym.spaces.discrete.Discrete", MultiDiscrete: "gym.spaces.multi_discrete.MultiDiscrete",
              Box: "gym.spaces.box.Box"}


# --------------------------------------------------------------------------- writing
def _pickle_by_reference(module, name, args=None):
    """Pickle `module.name` (or the call `module.name(*args)`) BY REFERENCE without
    importing the module: a throw-away stub stands in for it while pickling, so the
    bytes resolve to the real gym / SB3 class wherever those are installed."""
    saved = {k: sys.modules.get(k) for k in _parents(module)}
    try:
        for k in _parents(module):
            if k not in sys.modules or not hasattr(sys.modules[k], "__path__") and k != module:
                sys.modules[k] = types.ModuleType(k)
        mod = sys.modules[module] = types.ModuleType(module)
        cls = type(name, (), {"__module__": module, "__qualname__": name})
        setattr(mod, name, cls)
        if args is None:
            return pickle.dumps(cls, protocol=2)

        class _Call:
            def __reduce__(self):
                return cls, tuple(args)
        return pickle.dumps(_Call(), protocol=2)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _parents(module):
    parts = module.split(".")
    return [".".join(parts[:i + 1]) for i in range(len(parts))]


def _space_entry(space):
    """What SB3's data_to_json would write for a gym space."""
    mod, _, name = _GYM_TYPES[type(space)].rpartition(".")
    if isinstance(space, Discrete):
        blob = _pickle_by_reference(mod, name, (int(space.n),))
        fields = {"n": int(space.n), "_shape": [], "dtype": "int64", "_np_random": None}
    elif isinstance(space, MultiDiscrete):
        blob = _pickle_by_reference(mod, name, ([int(x) for x in space.nvec],))
        fields = {"nvec": str(np.asarray(space.nvec)), "_shape": [len(space.nvec)], "dtype": "int64",
                  "_np_random": None}
    else:
        low, high = np.asarray(space.low), np.asarray(space.high)
        dt = np.dtype(space.dtype)
        blob = _pickle_by_reference(mod, name, (low, high, tuple(low.shape), dt.type))
        with np.printoptions(threshold=100000):
            fields = {"dtype": dt.name, "_shape": list(low.shape), "low": str(low), "high": str(high),
                      "bounded_below": str(np.isfinite(low)), "bounded_above": str(np.isfinite(high)),
                      "_np_random": None}
    return {":type:": f"<class '{_GYM_TYPES[type(space)]}'>", ":serialized:": base64.b64encode(blob).decode(), **fields}


def _torch_bytes(obj):
    buf = io.BytesIO()
    torch.save(obj, buf)
    return buf.getvalue()


def optimizer_state_dict(names, adam_m, adam_v, adam_step, learning_rate, eps=1e-5):
    """torch.optim.Adam.state_dict() for SB3's single parameter group (Adam eps 1e-5,
    pantheonrl/algos/modular/policies.py:84-88); adam_m / adam_v are dicts name -> tensor."""
    state = {}
    if adam_step > 0:
        for i, n in enumerate(names):
            state[i] = {"step": torch.tensor(float(adam_step)), "exp_avg": adam_m[n].clone(),
                        "exp_avg_sq": adam_v[n].clone()}
    group = {"lr": float(learning_rate), "betas": (0.9, 0.999), "eps": eps, "weight_decay": 0, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "params": list(range(len(names)))}
    return {"state": state, "param_groups": [group]}


def save_zip(path, observation_space, action_space, hyper, policy_state, optimizer_state, counters=None):
    """Write an SB3-1.7.0-shaped PPO archive.  `path` gets ".zip" appended when it has
    no extension, like SB3's open_path."""
    path = str(path)
    if "." not in path.rsplit("/", 1)[-1]:
        path += ".zip"
    counters = counters or {}
    data = {
        "policy_class": {":type:": "<class 'abc.ABCMeta'>",
                         ":serialized:": base64.b64encode(_pickle_by_reference(
                             "stable_baselines3.common.policies", "ActorCriticPolicy")).decode(),
                         "__module__": "stable_baselines3.common.policies"},
        "observation_space": _space_entry(observation_space),
        "action_space": _space_entry(action_space),
        "num_timesteps": int(counters.get("num_timesteps", 0)),
        "_total_timesteps": int(counters.get("num_timesteps", 0)),
        "_num_timesteps_at_start": 0, "action_noise": None, "start_time": 0, "tensorboard_log": None,
        "_last_obs": None, "_last_episode_starts": None, "_last_original_obs": None, "_episode_num": 0,
        "use_sde": False, "sde_sample_freq": -1, "_current_progress_remaining": 0.0,
        "_n_updates": int(counters.get("n_updates", 0)),
        "clip_range_vf": None, "target_kl": None, "policy_kwargs": {},
        # extension fields (ignored by SB3's loader, used by ours)
        "b200": {"adam_step": int(counters.get("adam_step", 0)), "n_minibatches": int(hyper.get("n_minibatches", 0)),
                 "format": 1,
                 # constructor arguments SB3's PPO does not have (ADAP: context_size ...; ModularAlgorithm:
                 # num_partners ...) and counters beyond SB3's
                 "extra": {k: v for k, v in hyper.items() if k not in HYPER_KEYS and k != "n_minibatches"},
                 "counters": {k: v for k, v in counters.items() if k not in ("num_timesteps", "n_updates", "adam_step")}},
    }
    for k in HYPER_KEYS:
        if k in hyper:
            data[k] = hyper[k]
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps(data, indent=4))
        z.writestr("pytorch_variables.pth", _torch_bytes(None))
        z.writestr("policy.pth", _torch_bytes(dict(policy_state)))
        z.writestr("policy.optimizer.pth", _torch_bytes(optimizer_state))
        z.writestr("_stable_baselines3_version", SB3_VERSION)
        z.writestr("system_info.txt", f"pantheonrl_b200 checkpoint (SB3 {SB3_VERSION} archive layout)\n"
                                      f"PyTorch: {torch.__version__}\nNumpy: {np.__version__}\n")
    return path


# --------------------------------------------------------------------------- reading
def _parse_array(text, dtype):
    """numpy's str(array) -> array (SB3 stores space bounds / nvec this way)."""
    if "..." in text:
        raise ValueError("array was summarised with '...' when the archive was written")
    toks = re.findall(r"[-+]?(?:inf|nan|\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+)", text)
    return np.array([float(t) for t in toks]).astype(dtype)


def _parse_shape(v):
    if isinstance(v, (list, tuple)):
        return tuple(int(x) for x in v)
    return tuple(int(x) for x in re.findall(r"\d+", str(v)))


def space_from_entry(entry):
    """Rebuild a space from SB3's JSON rendering without touching the pickle."""
    t = entry.get(":type:", "")
    if "MultiDiscrete" in t:
        return MultiDiscrete(_parse_array(str(entry["nvec"]), np.int64))
    if "Discrete" in t:
        return Discrete(int(entry["n"]))
    if "Box" in t:
        shape = _parse_shape(entry["_shape"])
        dt = np.dtype(str(entry.get("dtype", "float32")))
        low = _parse_array(str(entry["low"]), dt).reshape(shape)
        high = _parse_array(str(entry["high"]), dt).reshape(shape)
        return Box(low, high, dtype=dt.type)
    raise ValueError(f"unsupported space in archive: {t!r}")


def load_zip(path):
    """-> dict(data, hyper, observation_space, action_space, policy, optimizer, counters)."""
    path = str(path)
    try:
        z = zipfile.ZipFile(path)
    except FileNotFoundError:
        z = zipfile.ZipFile(path + ".zip")
    with z:
        names = set(z.namelist())
        if "data" not in names or "policy.pth" not in names:
            raise ValueError(f"{path}: not an SB3-style archive (needs 'data' and 'policy.pth')")
        data = json.loads(z.read("data").decode())
        policy = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        optim = None
        if "policy.optimizer.pth" in names:
            optim = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), map_location="cpu", weights_only=True)
    hyper = {}
    for k in HYPER_KEYS:
        v = data.get(k)
        if isinstance(v, (int, float, bool)) or (k == "seed" and v is None and k in data):
            hyper[k] = v  # schedules (lr / clip_range given as callables) are pickled closures: keep defaults
    ext = data.get("b200", {})
    if ext.get("n_minibatches"):
        hyper["n_minibatches"] = int(ext["n_minibatches"])
    hyper.update(ext.get("extra", {}))
    adam_step = int(ext.get("adam_step", 0))
    if optim and optim.get("state") and not adam_step:
        st = next(iter(optim["state"].values()))
        adam_step = int(float(st["step"]))
    return {"data": data, "hyper": hyper,
            "observation_space": space_from_entry(data["observation_space"]),
            "action_space": space_from_entry(data["action_space"]),
            "policy": policy, "optimizer": optim,
            "counters": {"num_timesteps": int(data.get("num_timesteps", 0)), "n_updates": int(data.get("_n_updates", 0)),
                         "adam_step": adam_step, **ext.get("counters", {})}}


def adam_moments(names, optim):
    """Per-parameter exp_avg / exp_avg_sq dicts from an Adam state_dict (zeros when absent)."""
    m, v = {}, {}
    st = (optim or {}).get("state", {})
    for i, n in enumerate(names):
        s = st.get(i) or st.get(str(i))
        m[n] = s["exp_avg"] if s else None
        v[n] = s["exp_avg_sq"] if s else None
    return m, v
