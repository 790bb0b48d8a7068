"""GPU parity: fused policy forward + sample + log-prob + value vs the CPU
oracle, bit for bit (same fp32 evaluation order, same Philox draws)."""
import numpy as np
import pytest
import torch

import oracle
from pantheonrl_b200 import _lib, ops
from test_oracle_cpu import rand_liar_obs, rand_params

pytestmark = pytest.mark.gpu

SPACES = {
    "rps": (oracle.RPS_SPACE, lambda: _lib.Space.onehot([1], [3])),
    "liar": (oracle.LIAR_SPACE, lambda: _lib.Space.onehot(oracle.LIAR_NVEC, [7, 12])),
    "liar3": (oracle.LIAR3_SPACE, lambda: _lib.Space.onehot(oracle.LIAR_NVEC * 3, [7, 12])),  # frame stack: 90 slots
}


def _obs(name, B, seed=0):
    if name == "liar3":  # three stacked frames in a 96-byte row
        o = np.zeros((B, 96), np.uint8)
        for f in range(3):
            o[:, 30 * f:30 * f + 30] = rand_liar_obs(B, seed + 17 * f)[:, :30]
        return o
    return rand_liar_obs(B, seed) if name == "liar" else np.zeros((B, 32), np.uint8)


@pytest.mark.parametrize("name", ["rps", "liar", "liar3"])
@pytest.mark.parametrize("B", [1, 127, 128, 1000])
@pytest.mark.parametrize("scale", [0.05, 0.6])
def test_forward_sample_bit_exact(ctx, name, B, scale):
    okw, mk = SPACES[name]
    osp, gsp = oracle.make_space(**okw), mk()
    params = rand_params(osp, seed=B, scale=scale)
    obs = _obs(name, B, seed=B)
    want = oracle.policy_forward(osp, params, obs, seed=10, rng_stream=3, tick=77, slot=1, idx0=5)
    got = ops.policy_forward(gsp, torch.from_numpy(params).cuda(), torch.from_numpy(obs).cuda(),
                             seed=10, rng_stream=3, tick=77, slot=1, idx0=5)
    torch.cuda.synchronize()
    for k in ("logits", "value", "action", "logp", "entropy"):
        g = got[k].cpu().numpy()
        assert np.array_equal(g, want[k]), (k, np.abs(g.astype(np.float64) - want[k]).max())


@pytest.mark.parametrize("name", ["rps", "liar", "liar3"])
def test_evaluate_actions_bit_exact(ctx, name):
    okw, mk = SPACES[name]
    osp, gsp = oracle.make_space(**okw), mk()
    B = 513
    params = rand_params(osp, seed=3, scale=0.4)
    obs = _obs(name, B, seed=9)
    rng = np.random.RandomState(0)
    act = np.zeros((B, 4), np.uint8)
    for h, n in enumerate(okw["heads"]):
        act[:, h] = rng.randint(n, size=B)
    want = oracle.policy_forward(osp, params, obs, action_in=act)
    got = ops.policy_forward(gsp, torch.from_numpy(params).cuda(), torch.from_numpy(obs).cuda(),
                             action_in=torch.from_numpy(act).cuda())
    for k in ("value", "logp", "entropy"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    assert np.array_equal(got["action"].cpu().numpy(), act)


def test_device_math_matches_oracle_math(ctx):
    # tanh/exp/log are exercised through the policy: a 1-slot space whose first
    # layer is a lookup lets us push chosen pre-activations through pth_tanhf.
    osp = oracle.make_space(nvec=[200], heads=[3])
    gsp = _lib.Space.onehot([200], [3])
    P = oracle.param_count(osp)
    rng = np.random.RandomState(1)
    params = (rng.randn(P) * 0.5).astype(np.float32)
    params[:200 * 64] = np.linspace(-12, 12, 200 * 64).astype(np.float32)  # pi0.w rows = pre-activations
    obs = np.zeros((200, 32), np.uint8)
    obs[:, 0] = np.arange(200)
    want = oracle.policy_forward(osp, params, obs, seed=1)
    got = ops.policy_forward(gsp, torch.from_numpy(params).cuda(), torch.from_numpy(obs).cuda(), seed=1)
    for k in ("logits", "value", "logp", "entropy", "action"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k


def test_device_exp_log_tanh_bit_exact_dense_sweep(ctx):
    """Every fp32 value in a band of each function's domain (2^22 consecutive bit patterns per
    band) plus random values: device exp / log / tanh equal the oracle's bit for bit."""
    lib = _lib.load()

    def dev(which, x):
        dx = torch.from_numpy(x).cuda()
        dy = torch.empty_like(dx)
        _lib.check(lib.pth_debug_math(ctx.handle, which, dx.data_ptr(), dy.data_ptr(), dx.numel(),
                                      _lib.current_stream()), "pth_debug_math")
        return dy.cpu().numpy()

    def band(lo, n=1 << 22):
        start = np.float32(lo).view(np.uint32)
        return (start + np.arange(n, dtype=np.uint32)).view(np.float32)

    rng = np.random.RandomState(0)
    for lo in (0.62, 0.999, 2.0, 4.9, 9.99, 1e-3):  # around the branch points and the clamp at 10
        for sign in (1.0, -1.0):
            x = (band(lo) * np.float32(sign)).astype(np.float32)
            assert np.array_equal(dev(2, x), oracle.math_vec("tanh", x)), ("tanh", lo, sign)
    x = np.concatenate([rng.uniform(-12, 12, 1 << 21), rng.randn(1 << 20) * 1e-2, [0.0, -0.0, 10.0, -10.0, 50.0, 1e30,
                        -1e30, np.inf, -np.inf]]).astype(np.float32)
    assert np.array_equal(dev(2, x), oracle.math_vec("tanh", x))
    for lo in (-87.5, -20.0, -1.0, -1e-3, 1e-3, 0.69, 19.9, 87.9):
        x = band(abs(lo)) * np.float32(np.sign(lo))
        assert np.array_equal(dev(0, x), oracle.math_vec("exp", x)), ("exp", lo)
    x = rng.uniform(-100, 95, 1 << 21).astype(np.float32)
    assert np.array_equal(dev(0, x), oracle.math_vec("exp", x))
    for lo in (1.0, 1.41, 7.0, 1e3, 1e-30):
        x = band(lo)
        assert np.array_equal(dev(1, x), oracle.math_vec("log", x)), ("log", lo)


def test_sampling_is_a_pure_function_of_the_counter(ctx):
    gsp = _lib.Space.onehot([1], [3])
    osp = oracle.make_space(**oracle.RPS_SPACE)
    params = torch.from_numpy(rand_params(osp, seed=1, scale=1.0)).cuda()
    obs = torch.zeros(4096, 32, dtype=torch.uint8, device="cuda")
    a = ops.policy_forward(gsp, params, obs, seed=5, tick=9, idx0=0)["action"]
    b = ops.policy_forward(gsp, params, obs[:1000], seed=5, tick=9, idx0=3096)["action"]
    assert torch.equal(a[3096:], b)  # shards index one global stream
    c = ops.policy_forward(gsp, params, obs, seed=5, tick=10, idx0=0)["action"]
    assert not torch.equal(a, c)
