"""The alias modules of pantheonrl_b200.compat: the import lines of the reference's trainer.py /
examples resolve to this package, and — when /root/reference is present (authoring container) — the
reference's REAL trainer.py loads under them and its own generate_env / preset / input_check run."""
import argparse
import importlib.util
import os
import sys

import pytest

from pantheonrl_b200 import _lib, compat

sys.path.insert(0, os.path.dirname(__file__))
REF_TRAINER = "/root/reference/trainer.py"


@pytest.fixture
def aliases():
    done = compat.install()
    yield done
    compat.uninstall()


def test_reference_import_lines_resolve_to_this_package(aliases):
    assert aliases == ["gym", "overcookedgym", "pantheonrl", "stable_baselines3"]
    import gym
    from overcookedgym.overcooked_utils import LAYOUT_LIST
    from pantheonrl.algos.bc import BCShell, reconstruct_policy  # noqa: F401
    from pantheonrl.common.agents import OnPolicyAgent, StaticPolicyAgent  # noqa: F401
    from pantheonrl.common.multiagentenv import SimultaneousEnv, TurnBasedEnv  # noqa: F401
    from pantheonrl.common.wrappers import frame_wrap, recorder_wrap  # noqa: F401
    from pantheonrl.envs.blockworldgym import blockworld, simpleblockworld  # noqa: F401
    from pantheonrl.envs.liargym.liar import LiarDefaultAgent, LiarEnv  # noqa: F401
    from pantheonrl.envs.rpsgym.rps import RPSEnv, RPSWeightedAgent  # noqa: F401
    from stable_baselines3 import PPO
    from stable_baselines3.common.monitor import Monitor
    from stable_baselines3.common.vec_env import DummyVecEnv
    import pantheonrl_b200.common.agents as ours
    from pantheonrl_b200.ppo import PPO as OurPPO
    assert OnPolicyAgent is ours.OnPolicyAgent and PPO is OurPPO and "simple" in LAYOUT_LIST
    env = gym.make("RPS-v0")
    assert isinstance(env, RPSEnv) and env.getDummyEnv(1) is env
    assert compat.unwrap_env(DummyVecEnv([lambda: Monitor(env)])) is env
    with pytest.raises(NotImplementedError):
        simpleblockworld.SBWDefaultAgent()


def test_aliases_come_and_go():
    compat.install()
    compat.uninstall()
    assert not any(n.split(".")[0] in compat.ROOTS for n in sys.modules)


def _args():
    return argparse.Namespace(env="RPS-v0", ego="PPO", alt=["PPO"], env_config={}, ego_config={}, alt_config=None,
                              seed=10, device="auto", framestack=1, record=None, tensorboard_log=None,
                              tensorboard_name=None, ego_save=None, alt_save=None, verbose_partner=False,
                              share_latent=False, total_timesteps=500000)


@pytest.mark.skipif(not os.path.exists(REF_TRAINER), reason="the reference tree is only in the authoring container")
def test_the_reference_trainer_itself_loads_and_builds_its_env(aliases):
    """trainer.py of the reference, unmodified: every import resolves, preset(1) / input_check fill the
    arguments, generate_env builds the device-backed game through gym.make + getDummyEnv, and
    generate_ego reaches OUR PPO — which refuses to run without a GPU instead of falling back."""
    import torch
    spec = importlib.util.spec_from_file_location("reference_trainer", REF_TRAINER)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    args = ref.preset(_args(), 1)
    ref.input_check(args)
    assert args.alt_config == [{}] and args.ego_config == {"verbose": 1}
    assert args.ego_save == "models/RPS-v0-PPO-ego-10" and args.tensorboard_name == "RPS-v0-PPOPPO-10"
    env, altenv = ref.generate_env(args)
    from pantheonrl.envs.rpsgym.rps import RPSEnv
    assert isinstance(env, RPSEnv) and altenv is env
    # the same arguments through the trainer-shaped driver the GPU test uses
    import trainer_shaped as ts
    env2, alt2 = ts.make_envs(ts.default_args("RPS-v0", "PPO", ["PPO"], seed=10))
    assert type(env2) is type(env) and alt2 is env2
    assert ref.gen_default({"r": 2, "p": 1, "s": 1}, altenv).c0 == ts.make_partner(
        "DEFAULT", {"r": 2, "p": 1, "s": 1}, alt2, args, 0).c0 == 0.5
    if not torch.cuda.is_available():
        with pytest.raises(_lib.PthError):
            ref.generate_ego(env, args)
