"""Drop-in proof: driver code that only knows the REFERENCE's module names
(`gym`, `stable_baselines3`, `pantheonrl.*`, `overcookedgym.*`) trains through this package.

BASELINE configs[0] is `trainer.py RPS-v0 PPO PPO --seed 10`: ego and partner are built with the same
seed (trainer.py:111-112, 198-199) and therefore start from the same weights.  In the reference their
action samples still differ (one shared torch generator keeps advancing); here every PPO takes its own
Philox stream.  The test fails if the two learners ever share random numbers again: identical weights +
identical uniforms would make every RPS game a tie."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from pantheonrl_b200 import compat

sys.path.insert(0, os.path.dirname(__file__))
pytestmark = pytest.mark.gpu
REF_TRAINER = "/root/reference/trainer.py"


@pytest.fixture
def aliases():
    compat.install()
    yield
    compat.uninstall()


def test_trainer_flow_rps_ppo_vs_ppo_seed10(ctx, aliases, tmp_path):
    import trainer_shaped as ts
    args = ts.default_args("RPS-v0", "PPO", ["PPO"], seed=10, total_timesteps=2 * 2048,
                           record=str(tmp_path / "traj"), ego_save=str(tmp_path / "ego"),
                           alt_save=str(tmp_path / "alt"), ego_config={"verbose": 0})
    if os.path.exists(REF_TRAINER):  # never true on the GPU box; kept so the real file is used wherever it can be
        spec = importlib.util.spec_from_file_location("reference_trainer", REF_TRAINER)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        env, altenv = ref.generate_env(args)
        ego = ref.generate_ego(env, args)
        partners = ref.generate_partners(altenv, env, ego, args)
        ego.learn(total_timesteps=args.total_timesteps)
        ego.save(args.ego_save)
        partners[0].model.save(args.alt_save)
    else:
        env, ego, partners = ts.run(args)
    partner = partners[0]
    assert ego.policy.rng_stream != partner.model.policy.rng_stream
    tr = env.get_transitions()
    ea, aa = np.asarray(tr.egoacts).reshape(-1), np.asarray(tr.altacts).reshape(-1)
    assert ea.size == aa.size == 4096
    first_e, first_a = ea[:2048], aa[:2048]  # the first rollout: both policies still at their (equal) init
    assert not np.array_equal(first_e, first_a)
    non_tie = float((first_e != first_a).mean())
    assert 0.60 < non_tie < 0.73, non_tie  # two independent near-uniform players: 2/3
    for acts in (first_e, first_a):  # each stream is itself near uniform over {rock, paper, scissors}
        assert np.all(np.abs(np.bincount(acts.astype(np.int64), minlength=3) / 2048 - 1 / 3) < 0.05)
    # the ego's buffer holds the second rollout now; its rewards are the RPS payoffs of those steps
    r = ego.rollout_buffer.h["rewards"]
    want = ((ea[2048:] - aa[2048:] + 3) % 3).astype(np.float32)
    want[want == 2] = -1
    assert np.array_equal(r, want)
    assert ego._n_updates == 20 and partner.model._n_updates >= 10
    assert not torch.equal(ego.policy.params, partner.model.policy.params)
    from stable_baselines3 import PPO
    again = PPO.load(args.ego_save)
    assert torch.equal(again.policy.params, ego.policy.params)
    assert os.path.exists(args.alt_save + ".zip")


def test_overcooked_example_lines(ctx, aliases):
    """examples/overcookedtraining.py:8-30: the example's own statements."""
    import gym
    from stable_baselines3 import PPO

    from pantheonrl.common.agents import OnPolicyAgent
    from overcookedgym.overcooked_utils import LAYOUT_LIST

    layout = 'simple'
    assert layout in LAYOUT_LIST
    env = gym.make('OvercookedMultiEnv-v0', layout_name=layout)
    partner = OnPolicyAgent(PPO('MlpPolicy', env, verbose=1))
    env.add_partner_agent(partner)
    ego = PPO('MlpPolicy', env, verbose=1)
    ego.learn(total_timesteps=10000)

    assert ego.num_timesteps == 5 * 2048 and ego._n_updates == 50 and partner.model._n_updates == 40
    # unseeded learners do not share their initial weights (SB3 leaves the generators alone when seed=None)
    fresh_a, fresh_b = PPO('MlpPolicy', env), PPO('MlpPolicy', env)
    assert not torch.equal(fresh_a.policy.params, fresh_b.policy.params)
    assert bool(torch.isfinite(ego.policy.params).all())
