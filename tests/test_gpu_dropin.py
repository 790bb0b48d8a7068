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


def _run_adap(args, before_learn=None):
    import trainer_shaped as ts
    if os.path.exists(REF_TRAINER):
        spec = importlib.util.spec_from_file_location("reference_trainer", REF_TRAINER)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        env, altenv = ref.generate_env(args)
        ego = ref.generate_ego(env, args)
        partners = ref.generate_partners(altenv, env, ego, args)
    else:
        env, altenv = ts.make_envs(args)
        ego = ts.make_ego(env, args)
        partners = [ts.make_partner(k, c, altenv, args, i, ego) for i, (k, c) in enumerate(zip(args.alt, args.alt_config))]
        for p in partners:
            env.add_partner_agent(p)
    if before_learn is not None:
        before_learn(ego)
    ego.learn(total_timesteps=args.total_timesteps)
    return env, ego, partners


def test_trainer_flow_rps_adap_vs_adap(ctx, aliases, tmp_path):
    """`trainer.py RPS-v0 ADAP ADAP --seed 10` (trainer.py:127-128, 205-213): ADAP ego and an AdapAgent partner,
    each with its own context, resampled at every episode end (RPS: every step)."""
    import trainer_shaped as ts
    cfg = {"n_steps": 256, "batch_size": 64, "n_epochs": 3, "context_loss_coeff": 0.5}
    args = ts.default_args("RPS-v0", "ADAP", ["ADAP"], seed=10, total_timesteps=512,
                           ego_config=dict(cfg, verbose=0), alt_config=[dict(cfg)], ego_save=str(tmp_path / "ego"))
    env, ego, partners = _run_adap(args)
    partner = partners[0]
    p0 = type(ego)(env=env, seed=10, **cfg).policy.params
    assert not torch.equal(ego.policy.params, p0)  # it trained
    assert ego._n_updates == 6 and partner.model._n_updates >= 3
    for m in (ego, partner.model):
        cx = m.rollout_buffer.h["ctx"]
        filled = cx[:m.rollout_buffer.pos] if m is partner.model else cx
        assert filled.shape[1] == 3 and np.allclose(np.linalg.norm(filled, axis=1), 1.0, atol=1e-6)  # "l2" sampler
        assert len(np.unique(filled.round(5), axis=0)) > filled.shape[0] // 2  # a new context every episode
        cl = m.last_context_loss.cpu().numpy()
        assert cl.shape == (3 * 4,) and np.all((cl > 0) & (cl <= 1 + 1e-6))
        assert np.all(np.isfinite(m.last_stats.cpu().numpy()))
    assert not np.allclose(ego.rollout_buffer.h["ctx"][:8], partner.model.rollout_buffer.h["ctx"][:8])
    ego.save(args.ego_save)
    again = type(ego).load(args.ego_save)
    assert torch.equal(again.policy.params, ego.policy.params) and again.context_size == 3


def test_trainer_flow_liar_adap_share_latent(ctx, aliases):
    """`trainer.py LiarsDice-v0 ADAP ADAP --share-latent`: the partner copies the ego policy's context before every
    decision (trainer.py:210-213, adap/agent.py:74-75), so both buffers store the same contexts episode by episode."""
    import trainer_shaped as ts
    cfg = {"n_steps": 128, "batch_size": 64, "n_epochs": 2}
    args = ts.default_args("LiarsDice-v0", "ADAP", ["ADAP"], seed=3, total_timesteps=256, share_latent=True,
                           ego_config=dict(cfg, verbose=0), alt_config=[dict(cfg)])
    seen = set()

    def record_contexts(ego):
        seen.add(tuple(ego.policy.get_context().numpy().reshape(-1).round(6)))
        real = ego.policy.set_context

        def set_context(c):
            real(c)
            seen.add(tuple(ego.policy.get_context().numpy().reshape(-1).round(6)))
        ego.policy.set_context = set_context
    env, ego, partners = _run_adap(args, record_contexts)
    partner = partners[0]
    assert partner.latent_syncer is ego.policy
    alt_ctx = {tuple(r) for r in partner.model.rollout_buffer.h["ctx"][:partner.model.rollout_buffer.pos].round(6)}
    assert len(seen) > 10 and alt_ctx and alt_ctx <= seen  # every context the partner acted under was the ego's
    assert {tuple(r) for r in ego.rollout_buffer.h["ctx"].round(6)} <= seen
    assert ego._n_updates == 4


def test_trainer_flow_rps_modular_vs_two_ppo_partners(ctx, aliases, tmp_path):
    """`trainer.py RPS-v0 ModularAlgorithm PPO PPO --seed 10` (trainer.py:131-135): a ModularAlgorithm ego with one
    module per partner, two PPO partners; learn() plays n_steps with each partner in turn, then trains every module."""
    import trainer_shaped as ts
    cfg = {"n_steps": 128, "batch_size": 64, "n_epochs": 2}
    args = ts.default_args("RPS-v0", "ModularAlgorithm", ["PPO", "PPO"], seed=10, total_timesteps=2 * 2 * 128,
                           ego_config=dict(cfg, verbose=0, marginal_reg_coef=0.5), alt_config=[dict(cfg), dict(cfg)])
    env, ego, partners = _run_adap(args)
    assert ego.num_partners == 2 and len(ego.rollout_buffer) == 2
    assert ego.num_timesteps == 512 and ego._n_updates == 4
    assert ego.vf_steps == [8, 8] and ego.adam_step == 16      # 2 iterations x 2 epochs x 2 minibatches per partner
    for p in partners:                                          # each partner played (and recorded) its share
        assert p.num_timesteps == 256
    st, mg = ego.last_stats.cpu().numpy(), ego.last_marginal.cpu().numpy()
    assert st.shape == (8, 8) and np.all(np.isfinite(st)) and np.all(mg >= 0)
    path = str(tmp_path / "modular")
    ego.save(path)
    again = type(ego).load(path)
    assert torch.equal(again.policy.params, ego.policy.params) and again.num_partners == 2 and again.vf_steps == [8, 8]
    assert again.marginal_reg_coef == 0.5


def test_trainer_flow_liar_adap_mult_vs_adap_mult(ctx, aliases, tmp_path):
    """`trainer.py LiarsDice-v0 ADAP_MULT ADAP_MULT --seed 4` (trainer.py:129-130, 207-208): AdapPolicyMult on both sides."""
    import trainer_shaped as ts
    cfg = {"n_steps": 128, "batch_size": 64, "n_epochs": 2, "context_loss_coeff": 0.5}
    args = ts.default_args("LiarsDice-v0", "ADAP_MULT", ["ADAP_MULT"], seed=4, total_timesteps=256,
                           ego_config=dict(cfg, verbose=0), alt_config=[dict(cfg)])
    env, ego, partners = _run_adap(args)
    assert ego.mult and partners[0].model.mult and ego._n_updates == 4
    from pantheonrl_b200 import policy as pol
    assert ego.policy.params.numel() == pol.param_count_mult(ego.space, 3)
    cl = ego.last_context_loss.cpu().numpy()
    assert cl.shape == (4,) and np.all((cl > 0) & (cl <= 1 + 1e-6)) and np.all(np.isfinite(ego.last_stats.cpu().numpy()))
    path = str(tmp_path / "mult")
    ego.save(path)
    again = type(ego).load(path)
    assert again.mult and torch.equal(again.policy.params, ego.policy.params)
