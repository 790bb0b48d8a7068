"""Built-in envs whose rules also exist as device kernels (RPS-v0, LiarsDice-v0;
pantheonrl/envs/__init__.py:3-11) and OvercookedMultiEnv-v0 (overcookedgym/__init__.py:3-6)."""
from .rps import RPSEnv, RPSWeightedAgent  # noqa: F401
from .liar import LiarEnv, LiarDefaultAgent  # noqa: F401
from .overcooked import OvercookedMultiEnv  # noqa: F401

REGISTRY = {"RPS-v0": RPSEnv, "LiarsDice-v0": LiarEnv, "OvercookedMultiEnv-v0": OvercookedMultiEnv}


def make(env_id, **kwargs):
    return REGISTRY[env_id](**kwargs)
