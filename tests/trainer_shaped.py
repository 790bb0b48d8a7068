"""A driver with the call sequence of the reference's ``trainer.py`` (env -> ego -> partners ->
learn -> save; trainer.py:92-137, 182-228, 394-432), written against the REFERENCE's module
names only: ``gym``, ``stable_baselines3``, ``pantheonrl.*``.  It knows nothing about
pantheonrl_b200; tests run it after ``pantheonrl_b200.compat.install()``.

On the GPU box ``/root/reference`` does not exist, so the -m gpu drop-in test uses this file; in
the authoring container tests/test_compat_cpu.py loads the reference's real trainer.py under the
same aliases and checks that this driver and the real one agree on everything that runs
without a GPU (argument defaults, env construction)."""
import argparse


def default_args(env, ego, alt, **over):
    a = argparse.Namespace(env=env, ego=ego, alt=list(alt), total_timesteps=500000, device="auto", seed=None,
                           ego_config={}, alt_config=None, env_config={}, framestack=1, record=None,
                           ego_save=None, alt_save=None, share_latent=False, tensorboard_log=None,
                           tensorboard_name=None, verbose_partner=False, preset=None)
    for k, v in over.items():
        setattr(a, k, v)
    if a.alt_config is None:
        a.alt_config = [{} for _ in a.alt]
    a.ego_config.setdefault("verbose", 1)
    return a


def make_envs(args):
    import gym
    from pantheonrl.common.wrappers import frame_wrap, recorder_wrap
    env = gym.make(args.env, **args.env_config)
    altenv = env.getDummyEnv(1)
    if args.framestack > 1:
        env, altenv = frame_wrap(env, args.framestack), frame_wrap(altenv, args.framestack)
    if args.record is not None:
        env = recorder_wrap(env)
    return env, altenv


def make_ego(env, args):
    from stable_baselines3 import PPO
    kw = dict(args.ego_config, env=env, device=args.device, tensorboard_log=args.tensorboard_log)
    if args.seed is not None:
        kw["seed"] = args.seed
    if args.ego == "LOAD":
        from stable_baselines3.common.monitor import Monitor
        from stable_baselines3.common.vec_env import DummyVecEnv
        model = PPO.load(kw["location"])
        model.set_env(DummyVecEnv([lambda: Monitor(env)]))
        return model
    if args.ego in ("ADAP", "ADAP_MULT"):  # trainer.py:127-130
        from pantheonrl.algos.adap.adap_learn import ADAP
        from pantheonrl.algos.adap.policies import AdapPolicy, AdapPolicyMult
        return ADAP(policy=AdapPolicy if args.ego == "ADAP" else AdapPolicyMult, **kw)
    if args.ego == "ModularAlgorithm":  # trainer.py:131-135
        from pantheonrl.algos.modular.learn import ModularAlgorithm
        from pantheonrl.algos.modular.policies import ModularPolicy
        return ModularAlgorithm(policy=ModularPolicy, policy_kwargs=dict(num_partners=len(args.alt)), **kw)
    assert args.ego == "PPO"
    return PPO(policy="MlpPolicy", **kw)


def make_partner(kind, config, altenv, args, number, ego=None):
    from stable_baselines3 import PPO
    from pantheonrl.common.agents import OnPolicyAgent, StaticPolicyAgent
    from pantheonrl.envs.liargym.liar import LiarDefaultAgent, LiarEnv
    from pantheonrl.envs.rpsgym.rps import RPSEnv, RPSWeightedAgent
    if kind == "FIXED":
        return StaticPolicyAgent(PPO.load(config["location"]).policy)
    if kind == "DEFAULT":
        if isinstance(altenv, RPSEnv):
            return RPSWeightedAgent(**config)
        assert isinstance(altenv, LiarEnv) and not config
        return LiarDefaultAgent()
    agentarg = {}
    if args.tensorboard_log is not None:
        agentarg = {"tensorboard_log": args.tensorboard_log, "tb_log_name": f"{args.tensorboard_name}_alt_{number}"}
    config = dict(config, env=altenv, device=args.device, verbose=args.verbose_partner)
    if args.seed is not None:
        config["seed"] = args.seed
    if kind in ("ADAP", "ADAP_MULT"):  # trainer.py:205-213
        from pantheonrl.algos.adap.adap_learn import ADAP
        from pantheonrl.algos.adap.agent import AdapAgent
        from pantheonrl.algos.adap.policies import AdapPolicy, AdapPolicyMult
        shared = ego.policy if args.share_latent else None
        return AdapAgent(ADAP(policy=AdapPolicy if kind == "ADAP" else AdapPolicyMult, **config), latent_syncer=shared,
                         **agentarg)
    assert kind == "PPO"
    return OnPolicyAgent(PPO(policy="MlpPolicy", **config), **agentarg)


def run(args):
    """-> (env, ego, partners) after learn() and the optional saves."""
    env, altenv = make_envs(args)
    ego = make_ego(env, args)
    partners = []
    for i, (kind, cfg) in enumerate(zip(args.alt, args.alt_config)):
        p = make_partner(kind, cfg, altenv, args, i, ego)
        env.add_partner_agent(p)
        partners.append(p)
    learn = {"total_timesteps": args.total_timesteps}
    if args.tensorboard_log:
        learn["tb_log_name"] = args.tensorboard_name
    ego.learn(**learn)
    if args.record:
        env.get_transitions().write_transition(args.record)
    if args.ego_save:
        ego.save(args.ego_save)
    if args.alt_save:
        for i, p in enumerate(partners):
            if hasattr(p, "model"):
                p.model.save(args.alt_save if len(partners) == 1 else f"{args.alt_save}/{i}")
    return env, ego, partners
