"""Device-resident SB3-shaped MlpPolicy parameters.

Construction follows stable-baselines3 1.7.0 ActorCriticPolicy (SURVEY.md
Appendix A2; in-tree mirror pantheonrl/algos/modular/policies.py:111-118,
229-241): two separate 64-64 tanh towers, orthogonal init with gains sqrt(2)
(towers), 0.01 (action head), 1 (value head), zero biases, created in SB3's
order from ``torch.manual_seed(seed)`` so that an ego and a partner built with
the same seed start from identical weights (trainer.py:111-112, 198-199).

Initialisation is one-time host plumbing (torch CPU RNG + QR, exactly what the
reference does); the flat fp32 vector then lives on the device in the layout of
include/pantheon_b200.h (first-layer matrices input-major).
"""
import math

import numpy as np
import torch

from . import _lib

HID = 64


def _ortho(rows, cols, gain):
    w = torch.empty(rows, cols)
    torch.nn.init.orthogonal_(w, gain=gain)
    return w


def feature_dim(space, extra=0):
    """extra: AdapPolicy's context inputs behind the features (pantheonrl/algos/adap/policies.py:71-84)."""
    return (sum(space.nvec) if space.obs_kind == _lib.PTH_OBS_ONEHOT else space.obs_len) + extra


def param_count(space, extra=0):
    F, L = feature_dim(space, extra), sum(space.heads)
    return 2 * (HID * F + HID + HID * HID + HID) + L * HID + L + HID + 1


def tensor_shapes(space, extra=0):
    """(name, torch shape) in SB3 registration order."""
    F, L = feature_dim(space, extra), sum(space.heads)
    return [("mlp_extractor.policy_net.0.weight", (HID, F)), ("mlp_extractor.policy_net.0.bias", (HID,)),
            ("mlp_extractor.policy_net.2.weight", (HID, HID)), ("mlp_extractor.policy_net.2.bias", (HID,)),
            ("mlp_extractor.value_net.0.weight", (HID, F)), ("mlp_extractor.value_net.0.bias", (HID,)),
            ("mlp_extractor.value_net.2.weight", (HID, HID)), ("mlp_extractor.value_net.2.bias", (HID,)),
            ("action_net.weight", (L, HID)), ("action_net.bias", (L,)),
            ("value_net.weight", (1, HID)), ("value_net.bias", (1,))]


def init_flat(space, seed, extra=0):
    """SB3-style initial parameters as a flat float32 numpy vector (engine layout)."""
    F, L = feature_dim(space, extra), sum(space.heads)
    if seed is not None:
        torch.manual_seed(int(seed))
    # nn.Linear creation consumes RNG for the default init before orthogonal_
    # overwrites it; SB3 creates pi0, vf0, pi1, vf1, action_net, value_net.
    for fan_out, fan_in in ((HID, F), (HID, F), (HID, HID), (HID, HID), (L, HID), (1, HID)):
        torch.nn.Linear(fan_in, fan_out)
    g = math.sqrt(2)
    pi0, pi1 = _ortho(HID, F, g), _ortho(HID, HID, g)
    vf0, vf1 = _ortho(HID, F, g), _ortho(HID, HID, g)
    act = _ortho(L, HID, 0.01)
    val = _ortho(1, HID, 1.0)
    z = lambda n: torch.zeros(n)  # noqa: E731
    parts = [pi0.t().contiguous(), z(HID), pi1, z(HID), vf0.t().contiguous(), z(HID), vf1, z(HID),
             act, z(L), val, z(1)]
    return torch.cat([p.reshape(-1) for p in parts]).numpy().astype(np.float32)


def flat_to_state_dict(space, flat, extra=0):
    """Engine layout -> SB3 state_dict tensors (torch layout weight[out][in])."""
    flat = torch.as_tensor(np.asarray(flat, np.float32))
    out, o = {}, 0
    for i, (name, shape) in enumerate(tensor_shapes(space, extra)):
        n = int(np.prod(shape))
        chunk = flat[o:o + n]
        if i in (0, 4):
            chunk = chunk.reshape(shape[1], shape[0]).t().contiguous()
        out[name] = chunk.reshape(shape).clone()
        o += n
    return out


def state_dict_to_flat(space, sd, extra=0):
    parts = []
    for i, (name, shape) in enumerate(tensor_shapes(space, extra)):
        t = torch.as_tensor(sd[name]).float().reshape(shape)
        if i in (0, 4):
            t = t.t().contiguous()
        parts.append(t.reshape(-1))
    return torch.cat(parts).numpy().astype(np.float32)


# ---------------------------------------------------------------------------- AdapPolicyMult (adap/policies.py:134-283)
MULT_TRANSPOSED = (0, 6)  # the two first-layer matrices are stored input-major


def tensor_shapes_mult(space, C):
    """(name, torch shape) of an AdapPolicyMult in flat order: per tower first layer | scaling 64 -> 64 C | second layer."""
    F, L = feature_dim(space), sum(space.heads)
    out = []
    for tower in ("agent", "value"):
        out += [(f"mlp_extractor.{tower}_branch_1.0.weight", (HID, F)), (f"mlp_extractor.{tower}_branch_1.0.bias", (HID,)),
                (f"mlp_extractor.{tower}_scaling.0.weight", (HID * C, HID)), (f"mlp_extractor.{tower}_scaling.0.bias", (HID * C,)),
                (f"mlp_extractor.{tower}_branch_2.0.weight", (HID, HID)), (f"mlp_extractor.{tower}_branch_2.0.bias", (HID,))]
    return out + [("action_net.weight", (L, HID)), ("action_net.bias", (L,)), ("value_net.weight", (1, HID)),
                  ("value_net.bias", (1,))]


def param_count_mult(space, C):
    return sum(int(np.prod(s)) for _, s in tensor_shapes_mult(space, C))


def init_flat_mult(space, seed, C):
    """MultModel.__init__ creates pi0, vf0, pi1, vf1 (zip_longest), then the two scaling layers, then SB3 creates the
    heads; ActorCriticPolicy._build orthogonalises module by module (mlp_extractor = the MultModel, gain sqrt 2, in
    registration order agent_branch_1, agent_scaling, agent_branch_2, value_branch_1, value_scaling, value_branch_2)."""
    F, L = feature_dim(space), sum(space.heads)
    if seed is not None:
        torch.manual_seed(int(seed))
    for fan_out, fan_in in ((HID, F), (HID, F), (HID, HID), (HID, HID), (HID * C, HID), (HID * C, HID), (L, HID), (1, HID)):
        torch.nn.Linear(fan_in, fan_out)
    g = math.sqrt(2)
    z = lambda n: torch.zeros(n)  # noqa: E731
    a1, as_, a2 = _ortho(HID, F, g), _ortho(HID * C, HID, g), _ortho(HID, HID, g)
    v1, vs, v2 = _ortho(HID, F, g), _ortho(HID * C, HID, g), _ortho(HID, HID, g)
    parts = [a1.t().contiguous(), z(HID), as_, z(HID * C), a2, z(HID), v1.t().contiguous(), z(HID), vs, z(HID * C), v2,
             z(HID), _ortho(L, HID, 0.01), z(L), _ortho(1, HID, 1.0), z(1)]
    return torch.cat([p.reshape(-1) for p in parts]).numpy().astype(np.float32)


def flat_to_state_dict_mult(space, flat, C):
    flat = torch.as_tensor(np.asarray(flat, np.float32))
    out, o = {}, 0
    for i, (name, shape) in enumerate(tensor_shapes_mult(space, C)):
        n = int(np.prod(shape))
        chunk = flat[o:o + n]
        if i in MULT_TRANSPOSED:
            chunk = chunk.reshape(shape[1], shape[0]).t().contiguous()
        out[name] = chunk.reshape(shape).clone()
        o += n
    return out


def state_dict_to_flat_mult(space, sd, C):
    parts = []
    for i, (name, shape) in enumerate(tensor_shapes_mult(space, C)):
        t = torch.as_tensor(sd[name]).float().reshape(shape)
        parts.append((t.t().contiguous() if i in MULT_TRANSPOSED else t).reshape(-1))
    return torch.cat(parts).numpy().astype(np.float32)
