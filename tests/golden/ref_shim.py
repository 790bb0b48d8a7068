"""Import the reference's env / routing code verbatim from /root/reference.

Test-time tooling only (used by make_golden.py in the authoring container;
/root/reference does not exist on the GPU box).  gym and stable_baselines3 are
not installed, so minimal in-memory stub modules stand in for the names the
reference imports at module scope; none of the stubbed functionality is
executed by RPSEnv / LiarEnv / MultiAgentEnv.
"""
import sys
import types

import numpy as np

REF = "/root/reference"


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install():
    if "pantheonrl" in sys.modules and getattr(sys.modules["pantheonrl"], "__file__", "").startswith(REF):
        return
    gym = _mod("gym")
    spaces = _mod("gym.spaces")
    gym.spaces = spaces

    class Env:
        pass

    class Space:
        pass

    class Discrete(Space):
        def __init__(self, n):
            self.n, self.shape = n, ()

    class MultiDiscrete(Space):
        def __init__(self, nvec):
            self.nvec = np.array(nvec)
            self.shape = (len(nvec),)

    class MultiBinary(Space):
        def __init__(self, n):
            self.n, self.shape = n, (n,)

    class Box(Space):
        def __init__(self, low, high, dtype=np.float32):
            self.low, self.high, self.dtype, self.shape = low, high, dtype, np.shape(low)

    gym.Env = Env
    for c in (Space, Discrete, MultiDiscrete, MultiBinary, Box):
        setattr(spaces, c.__name__, c)
    envs = _mod("gym.envs")
    reg = _mod("gym.envs.registration")
    gym.envs = envs
    envs.registration = reg
    reg.register = lambda id=None, entry_point=None, **kw: None

    _mod("stable_baselines3")
    for sub in ("common", "common.utils", "common.policies", "common.on_policy_algorithm",
                "common.off_policy_algorithm", "common.base_class"):
        _mod("stable_baselines3." + sub)
    u = sys.modules["stable_baselines3.common.utils"]
    u.configure_logger = u.should_collect_more_steps = u.safe_mean = u.obs_as_tensor = None
    sys.modules["stable_baselines3.common.policies"].ActorCriticPolicy = object
    sys.modules["stable_baselines3.common"].policies = sys.modules["stable_baselines3.common.policies"]
    sys.modules["stable_baselines3.common.on_policy_algorithm"].OnPolicyAlgorithm = object
    sys.modules["stable_baselines3.common.off_policy_algorithm"].OffPolicyAlgorithm = object
    sys.modules["stable_baselines3.common.base_class"].BaseAlgorithm = object
    if REF not in sys.path:
        sys.path.insert(0, REF)
