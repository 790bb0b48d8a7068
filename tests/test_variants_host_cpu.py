"""Host-side logic of the ADAP / ModularAlgorithm facades that needs no GPU: parameter layouts agree with the C ABI and
the oracle, state-dict round trips, the row format of the ADAP buffer, the reference-RNG samplers, alias modules."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from oracle import update as oupd
from pantheonrl_b200 import _lib, adap, compat, modular, policy as pol
from pantheonrl_b200.ppo import HostStagedBuffer

LIAR = _lib.Space.onehot(oracle.LIAR_NVEC, [7, 12])


def test_param_counts_agree_between_facade_abi_and_oracle():
    lib = _lib.load()
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    for C_ in (1, 3, 8):
        assert pol.param_count(LIAR, C_) == lib.pth_adap_param_count(C.byref(LIAR), C_) == oracle.adap_param_count(osp, C_)
        assert pol.init_flat(LIAR, 3, C_).size == pol.param_count(LIAR, C_)
    for Pn in (1, 2, 8):
        n = lib.pth_modular_param_count(C.byref(LIAR), Pn)
        assert n == oupd.modular_param_count(osp, Pn) == modular.init_flat(LIAR, 3, Pn).size
    assert lib.pth_modular_param_count(C.byref(LIAR), 9) < 0 and lib.pth_adap_param_count(C.byref(LIAR), 9) < 0


def test_modular_init_follows_the_reference_construction_order():
    """Same seed -> the main network of a ModularPolicy differs from a plain MlpPolicy's (the partner modules'
    default inits are drawn before the single orthogonal pass, modular/policies.py:229-270), heads keep their gains."""
    a, b = modular.init_flat(LIAR, 5, 2), pol.init_flat(LIAR, 5)
    assert not np.array_equal(a[:b.size], b)
    L, main = 19, pol.param_count(LIAR)
    blk = (a.size - main) // 2
    w_act = a[main + 4 * (64 * 64 + 64):main + 4 * (64 * 64 + 64) + L * 64].reshape(L, 64)
    assert np.allclose(w_act @ w_act.T, 1e-4 * np.eye(L), atol=1e-6)  # orthogonal rows, gain 0.01
    w0 = a[main:main + 64 * 64].reshape(64, 64)
    assert np.allclose(w0 @ w0.T, 2 * np.eye(64), atol=1e-4)          # gain sqrt(2)
    assert not a[main + 64 * 64:main + 64 * 64 + 64].any() and blk == 17940


def test_modular_state_dict_round_trip_and_names():
    class P(modular.ModularDevicePolicy):
        def __init__(self):  # no device
            self.space, self.num_partners = LIAR, 2
    p = P()
    flat = np.random.RandomState(0).randn(modular.init_flat(LIAR, 1, 2).size).astype(np.float32)
    sd = p.flat_to_dict(flat)
    assert "partner_mlp_extractor.1.value_net.2.weight" in sd and "partner_action_net.0.bias" in sd
    assert sd["mlp_extractor.policy_net.0.weight"].shape == (64, 270) and sd["partner_value_net.1.weight"].shape == (1, 64)
    assert np.array_equal(p.dict_to_flat(sd), flat)


def test_adap_buffer_splits_observation_and_context():
    buf = adap.AdapBuffer(4, "cpu", 0.99, 0.95, box=False, row=32, context_size=3)
    row = np.concatenate([np.arange(30) % 7, [0.25, -0.5, 0.75]]).astype(np.float64)
    buf.add(row, np.array([[3, 4]]), 1.0, True, torch.tensor([0.5]), torch.tensor([-1.0]))
    assert np.array_equal(buf.h["obs"][0, :30], (np.arange(30) % 7).astype(np.uint8)) and not buf.h["obs"][0, 30:].any()
    assert np.array_equal(buf.h["ctx"][0], np.float32([0.25, -0.5, 0.75])) and buf.pos == 1
    assert isinstance(buf, HostStagedBuffer)


@pytest.mark.parametrize("name,C_", [("l2", 3), ("unit_square", 4), ("positive_square", 2), ("categorical", 5),
                                     ("natural_numbers", 1)])
def test_reference_rng_samplers_follow_the_reference(name, C_):
    """SAMPLERS[name](ctx_size, num, torch=True): same draws as pantheonrl/algos/adap/util.py:42-94 from torch's generator."""
    torch.manual_seed(4)
    got = adap.SAMPLERS[name](ctx_size=C_, num=6, torch=True)
    torch.manual_seed(4)
    if name == "l2":
        c = torch.rand(6, C_) * 2 - 1
        want = c / (torch.sum(c ** 2, dim=-1).reshape(6, 1)) ** (1 / 2)
    elif name == "unit_square":
        want = torch.rand(6, C_) * 2 - 1
    elif name == "positive_square":
        want = torch.rand(6, C_)
    elif name == "categorical":
        want = torch.zeros(6, C_)
        want[torch.arange(6), torch.randint(0, C_, size=(6,))] = 1
    else:
        want = torch.randint(0, C_, size=(6, 1)).float()
    assert torch.equal(got, want)


def test_alias_modules_resolve_the_variants():
    compat.install()
    try:
        from pantheonrl.algos.adap.adap_learn import ADAP
        from pantheonrl.algos.adap.agent import AdapAgent
        from pantheonrl.algos.adap.policies import AdapPolicy, AdapPolicyMult
        from pantheonrl.algos.adap.util import SAMPLERS
        from pantheonrl.algos.modular.learn import ModularAlgorithm
        from pantheonrl.algos.modular.policies import ModularPolicy
        assert ADAP is adap.ADAP and AdapAgent is adap.AdapAgent and AdapPolicy is adap.AdapPolicy
        assert ModularAlgorithm is modular.ModularAlgorithm and ModularPolicy is modular.ModularPolicy
        assert set(SAMPLERS) == {"l2", "unit_square", "positive_square", "categorical", "natural_numbers"}
        assert AdapPolicyMult is adap.AdapPolicyMult
    finally:
        compat.uninstall()


def test_adap_mult_layout_agrees_between_facade_abi_and_oracle():
    lib = _lib.load()
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    for C_ in (1, 3, 8):
        n = pol.param_count_mult(LIAR, C_)
        assert n == lib.pth_adap_mult_param_count(C.byref(LIAR), C_) == oracle.adap_mult_param_count(osp, C_)
        flat = pol.init_flat_mult(LIAR, 3, C_)
        assert flat.size == n
        sd = pol.flat_to_state_dict_mult(LIAR, flat, C_)
        assert sd["mlp_extractor.agent_scaling.0.weight"].shape == (64 * C_, 64)
        assert np.array_equal(pol.state_dict_to_flat_mult(LIAR, sd, C_), flat)
    # the scaling layer is orthogonal with gain sqrt(2) (64 C x 64: orthonormal columns)
    ws = pol.flat_to_state_dict_mult(LIAR, pol.init_flat_mult(LIAR, 3, 3), 3)["mlp_extractor.value_scaling.0.weight"].numpy()
    assert np.allclose(ws.T @ ws, 2 * np.eye(64), atol=1e-4)


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/pantheonrl"), reason="needs the reference tree")
def test_reference_adap_agent_cannot_record_documented_defect():
    """DESIGN.md 9 claims the reference's own AdapAgent.get_action (adap/agent.py:121-127) raises whenever it
    records: it reshapes the (observation ++ context) row to the policy's plain observation shape.  Run it."""
    import subprocess
    import sys
    code = r'''
import sys, types
sys.path.insert(0, "tests/golden")
import ref_shim; ref_shim.install()
import numpy as np, torch as th
def stub(name, **attrs):
    m = sys.modules.get(name) or types.ModuleType(name); sys.modules[name] = m
    for k, v in attrs.items(): setattr(m, k, v)
class Lg:
    def record(self, *a, **k): pass
    def dump(self, *a, **k): pass
stub("stable_baselines3.common.utils", configure_logger=lambda *a, **k: Lg(), safe_mean=np.mean,
     should_collect_more_steps=None, obs_as_tensor=None)
stub("pantheonrl.algos.adap.adap_learn", ADAP=object)
stub("pantheonrl.algos.adap.policies", AdapPolicy=object)
stub("pantheonrl.algos.adap.util", SAMPLERS={"l2": lambda ctx_size, num, torch: th.ones(1, ctx_size)})
from pantheonrl.algos.adap import agent as A
A.action_from_policy = lambda obs, policy: (np.array([[1, 2]]), th.tensor([0.5]), th.tensor([-1.0]))
A.resample_noise = lambda m, n: None
A.clip_actions = lambda a, m: a
class Buf:
    obs_shape = (30,)
    def reset(self): pass
    def add(self, *a): pass
class Pol:
    observation_space = types.SimpleNamespace(shape=(30,)); action_space = types.SimpleNamespace(shape=(2,))
    def get_context(self): return th.ones(1, 3)
    def set_context(self, c): pass
class Model:
    verbose = 0; context_size = 3; n_steps = 100; context_sampler = "l2"; rollout_buffer = Buf(); policy = Pol()
    def set_logger(self, l): pass
from pantheonrl.common.observation import Observation
ag = A.AdapAgent(Model())
ag.get_action(Observation(np.zeros(30)), record=False)      # not recording: fine
try:
    ag.get_action(Observation(np.zeros(30)), record=True)
    print("RECORDED")
except ValueError as e:
    print("VALUEERROR", e)
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       cwd=__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    assert "VALUEERROR cannot reshape array of size 33 into shape (1,30)" in r.stdout, r.stdout + r.stderr
