"""Module aliases so that driver code written for the reference runs unmodified.

``trainer.py`` and ``examples/overcookedtraining.py`` of the reference begin with

    import gym
    from stable_baselines3 import PPO
    from stable_baselines3.common.vec_env import DummyVecEnv
    from stable_baselines3.common.monitor import Monitor
    from pantheonrl.common.agents import OnPolicyAgent, StaticPolicyAgent
    from pantheonrl.envs.rpsgym.rps import RPSEnv, RPSWeightedAgent
    from overcookedgym.overcooked_utils import LAYOUT_LIST
    ...

(trainer.py:1-27, examples/overcookedtraining.py:8-12).  ``install()`` registers those module
names in ``sys.modules`` and points them at this package, so the same lines import the B200
path: ``PPO`` is ``pantheonrl_b200.ppo.PPO``, ``OnPolicyAgent`` is
``pantheonrl_b200.common.agents.OnPolicyAgent``, ``gym.make`` builds the device-backed envs.
A name that already resolves to a real installation (a genuine ``gym`` / ``stable_baselines3``
/ ``pantheonrl``) is left alone unless ``force=True``.

The pieces of the reference that are outside this package's scope (blockworld games,
PettingZoo adapter) import as placeholders that raise on use, so
``import``-time references in trainer.py resolve and the failure is loud and late.
"""
import importlib
import importlib.util
import sys
import types


class Monitor:
    """stable_baselines3.common.monitor.Monitor as trainer.py:119 uses it: a transparent
    wrapper (episode statistics are kept by PPO.learn itself, ppo.collect_rollouts_single_env)."""

    def __init__(self, env, filename=None, **_):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)


class DummyVecEnv:
    """stable_baselines3.common.vec_env.DummyVecEnv([lambda: env]) (trainer.py:119): holds the
    one env; PPO.set_env unwraps it."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        if self.num_envs != 1:
            raise ValueError("host-side vectorisation is not how this package scales: pass n_envs= to PPO")

    def __getattr__(self, name):
        return getattr(self.envs[0], name)


def unwrap_env(env):
    """The MultiAgentEnv inside DummyVecEnv([lambda: Monitor(env)])."""
    while True:
        if isinstance(env, DummyVecEnv):
            env = env.envs[0]
        elif isinstance(env, Monitor):
            env = env.env
        else:
            return env


class _OutOfScope:
    """Placeholder for a reference class this package does not rebuild."""
    _what = "this component"

    def __init__(self, *a, **k):
        raise NotImplementedError(f"{self._what} is outside the scope of pantheonrl_b200 (SURVEY.md 2.2)")


def _placeholder(name, what):
    return type(name, (_OutOfScope,), {"_what": what})


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__pantheonrl_b200_alias__ = True
    return m


def _tree():
    """name -> module for every alias; parents are created as packages."""
    from . import bc, envs, ppo, spaces
    from .common import agents, multiagentenv, observation, trajsaver, wrappers
    from .envs import liar, overcooked, rps

    def gym_make(env_id, **kwargs):
        return envs.make(env_id, **kwargs)

    gym_spaces = _module("gym.spaces", Space=spaces.Space, Discrete=spaces.Discrete,
                         MultiDiscrete=spaces.MultiDiscrete, Box=spaces.Box)
    sbw = _module("pantheonrl.envs.blockworldgym.simpleblockworld",
                  PartnerEnv=_placeholder("PartnerEnv", "BlockEnv-v0"),
                  SBWDefaultAgent=_placeholder("SBWDefaultAgent", "BlockEnv-v0"))
    bw = _module("pantheonrl.envs.blockworldgym.blockworld",
                 PartnerEnv=_placeholder("PartnerEnv", "BlockEnv-v1"),
                 DefaultConstructorAgent=_placeholder("DefaultConstructorAgent", "BlockEnv-v1"))
    mods = {
        "gym": _module("gym", make=gym_make, spaces=gym_spaces, Env=object),
        "gym.spaces": gym_spaces,
        "stable_baselines3": _module("stable_baselines3", PPO=ppo.PPO),
        "stable_baselines3.common": _module("stable_baselines3.common"),
        "stable_baselines3.common.vec_env": _module("stable_baselines3.common.vec_env", DummyVecEnv=DummyVecEnv),
        "stable_baselines3.common.monitor": _module("stable_baselines3.common.monitor", Monitor=Monitor),
        "pantheonrl": _module("pantheonrl"),
        "pantheonrl.common": _module("pantheonrl.common"),
        "pantheonrl.common.agents": agents,
        "pantheonrl.common.multiagentenv": multiagentenv,
        "pantheonrl.common.observation": observation,
        "pantheonrl.common.trajsaver": trajsaver,
        "pantheonrl.common.wrappers": wrappers,
        "pantheonrl.algos": _module("pantheonrl.algos"),
        "pantheonrl.algos.bc": bc,
        "pantheonrl.envs": _module("pantheonrl.envs"),
        "pantheonrl.envs.rpsgym": _module("pantheonrl.envs.rpsgym"),
        "pantheonrl.envs.rpsgym.rps": rps,
        "pantheonrl.envs.liargym": _module("pantheonrl.envs.liargym"),
        "pantheonrl.envs.liargym.liar": liar,
        "pantheonrl.envs.blockworldgym": _module("pantheonrl.envs.blockworldgym", simpleblockworld=sbw,
                                                 blockworld=bw),
        "pantheonrl.envs.blockworldgym.simpleblockworld": sbw,
        "pantheonrl.envs.blockworldgym.blockworld": bw,
        "overcookedgym": _module("overcookedgym"),
        "overcookedgym.overcooked": overcooked,
        "overcookedgym.overcooked_utils": _module("overcookedgym.overcooked_utils",
                                                  LAYOUT_LIST=list(overcooked.LAYOUT_LIST),
                                                  NAME_TRANSLATION=dict(overcooked.NAME_TRANSLATION)),
    }
    # pantheonrl.algos.adap / .modular (trainer.py:14-19)
    from . import adap as adap_mod, modular as modular_mod
    mods.update({
        "pantheonrl.algos.adap": _module("pantheonrl.algos.adap"),
        "pantheonrl.algos.adap.adap_learn": _module("pantheonrl.algos.adap.adap_learn", ADAP=adap_mod.ADAP),
        "pantheonrl.algos.adap.policies": _module("pantheonrl.algos.adap.policies", AdapPolicy=adap_mod.AdapPolicy,
                                                  AdapPolicyMult=adap_mod.AdapPolicyMult),
        "pantheonrl.algos.adap.agent": _module("pantheonrl.algos.adap.agent", AdapAgent=adap_mod.AdapAgent),
        "pantheonrl.algos.adap.util": _module("pantheonrl.algos.adap.util", SAMPLERS=adap_mod.SAMPLERS),
        "pantheonrl.algos.modular": _module("pantheonrl.algos.modular"),
        "pantheonrl.algos.modular.learn": _module("pantheonrl.algos.modular.learn",
                                                  ModularAlgorithm=modular_mod.ModularAlgorithm),
        "pantheonrl.algos.modular.policies": _module("pantheonrl.algos.modular.policies",
                                                     ModularPolicy=modular_mod.ModularPolicy),
    })
    return mods


ROOTS = ("gym", "stable_baselines3", "pantheonrl", "overcookedgym")
_installed = []


def _real(root):
    if root in {n.split(".")[0] for n in _installed}:
        return False
    if root in sys.modules:
        return True
    try:
        return importlib.util.find_spec(root) is not None
    except (ImportError, ValueError):
        return False


def install(force=False):
    """Register the aliases.  Returns the root names that now point at this package."""
    taken = [r for r in ROOTS if _real(r) and not force]
    mods = _tree()
    done = set()
    for name, m in mods.items():
        root = name.split(".")[0]
        if root in taken:
            continue
        sys.modules[name] = m
        _installed.append(name)
        done.add(root)
        if "." in name:  # `import a.b` must also find b as an attribute of a
            parent, leaf = name.rsplit(".", 1)
            if getattr(mods.get(parent), "__pantheonrl_b200_alias__", False) and not hasattr(mods[parent], leaf):
                setattr(mods[parent], leaf, m)
    return sorted(done)


def uninstall():
    """Remove what install() registered."""
    while _installed:
        sys.modules.pop(_installed.pop(), None)
