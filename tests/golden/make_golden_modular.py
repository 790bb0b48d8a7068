"""Pin the ModularAlgorithm half of the oracle on the reference's OWN code (authoring container only:
python tests/golden/make_golden_modular.py).

Executed verbatim from /root/reference (stable-baselines3 itself is not installable here, so SB3's base classes
are stand-ins; every line that computes something on this path is the reference's):

  * ModularAlgorithm.train      pantheonrl/algos/modular/learn.py:221-351, unbound on a duck-typed `self`
                                (PPO's losses per partner buffer + the marginal regulariser :298-318)
  * ModularPolicy._get_latent / _get_action_dist_from_latent / evaluate_actions / get_action_logits_from_obs / forward
                                pantheonrl/algos/modular/policies.py:273-396, bound to the torch modules of
                                oracle/sb3_torch.ModularMlpPolicy (built like `_build` + `do_init_weights` :229-270)

torch here is 2.x: `optimizer.zero_grad()` sets gradients to None, so the value modules of the partners that are
not being trained are skipped by Adam (the reference's pinned torch 1.13.1 zero-fills them after their first use
instead).  Writes modular.npz.
"""
import os
import sys
import types

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_shim  # noqa: E402

ref_shim.install()
th.set_num_threads(1)  # bit-reproducible fixtures: torch's CPU GEMM blocking depends on the thread count


def stub(name, **attrs):
    m = sys.modules.get(name) or types.ModuleType(name)
    sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


class Distribution:
    pass


class CategoricalDistribution(Distribution):
    """SB3 1.7.0 distributions.CategoricalDistribution (the methods this path calls)."""

    def proba_distribution(self, action_logits):
        self.distribution = th.distributions.Categorical(logits=action_logits)
        return self

    def log_prob(self, actions):
        return self.distribution.log_prob(actions)

    def entropy(self):
        return self.distribution.entropy()

    def get_actions(self, deterministic=False):
        return th.argmax(self.distribution.probs, dim=1) if deterministic else self.distribution.sample()


class MultiCategoricalDistribution(Distribution):
    """SB3 1.7.0 distributions.MultiCategoricalDistribution."""

    def __init__(self, action_dims):
        self.action_dims = action_dims

    def proba_distribution(self, action_logits):
        self.distribution = [th.distributions.Categorical(logits=split)
                             for split in th.split(action_logits, tuple(self.action_dims), dim=1)]
        return self

    def log_prob(self, actions):
        return th.stack([dist.log_prob(action) for dist, action in zip(self.distribution, th.unbind(actions, dim=1))],
                        dim=1).sum(dim=1)

    def entropy(self):
        return th.stack([dist.entropy() for dist in self.distribution], dim=1).sum(dim=1)

    def get_actions(self, deterministic=False):
        return th.stack([dist.sample() for dist in self.distribution], dim=1)


class _Never(Distribution):
    pass


stub("stable_baselines3", PPO=object)
dist_mod = stub("stable_baselines3.common.distributions", Distribution=Distribution,
                CategoricalDistribution=CategoricalDistribution, MultiCategoricalDistribution=MultiCategoricalDistribution,
                DiagGaussianDistribution=_Never, BernoulliDistribution=type("B", (_Never,), {}),
                StateDependentNoiseDistribution=type("S", (_Never,), {}), make_proba_distribution=None)
common = sys.modules["stable_baselines3.common"]
common.distributions = dist_mod
common.logger = stub("stable_baselines3.common.logger")
stub("stable_baselines3.common.preprocessing", preprocess_obs=None, is_image_space=None, get_action_dim=None)
stub("stable_baselines3.common.torch_layers", FlattenExtractor=object, BaseFeaturesExtractor=object, create_mlp=None,
     NatureCNN=object, MlpExtractor=object)
stub("stable_baselines3.common.utils", get_device=None, is_vectorized_observation=None, safe_mean=None,
     explained_variance=None, get_schedule_fn=lambda v: (lambda _: v))
stub("stable_baselines3.common.vec_env", VecTransposeImage=object, VecEnv=object)
stub("stable_baselines3.common.policies", BasePolicy=object, ActorCriticPolicy=object)
stub("stable_baselines3.common.on_policy_algorithm", OnPolicyAlgorithm=object)
stub("stable_baselines3.common.base_class", BaseAlgorithm=object)
stub("stable_baselines3.common.buffers", RolloutBuffer=object)
stub("stable_baselines3.common.callbacks", BaseCallback=object)
stub("stable_baselines3.common.type_aliases", GymEnv=object, MaybeCallback=object, Schedule=object)
from pantheonrl.algos.modular import learn as modular_learn  # noqa: E402  (the reference's files, verbatim)
from pantheonrl.algos.modular.policies import ModularPolicy  # noqa: E402

import oracle  # noqa: E402
from oracle import sb3_torch  # noqa: E402
from oracle import update as oupd  # noqa: E402
from test_oracle_update import make_batch  # noqa: E402


class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Buffer:
    """RolloutBuffer.get as SB3 defines it: consecutive slices of one permutation per epoch."""

    def __init__(self, obs, act, old_logp, adv, ret, old_values, perms):
        self.t = dict(observations=th.as_tensor(obs), actions=th.as_tensor(act).float(),
                      old_log_prob=th.as_tensor(old_logp), advantages=th.as_tensor(adv), returns=th.as_tensor(ret),
                      old_values=th.as_tensor(old_values))
        self.perms, self.epoch = perms, 0

    def get(self, batch_size):
        perm = np.asarray(self.perms[self.epoch])
        self.epoch += 1
        for s in range(0, len(perm), batch_size):
            idx = perm[s:s + batch_size]
            yield Data(**{k: v[idx] for k, v in self.t.items()})


class Log:
    def __init__(self):
        self.kv = {}

    def record(self, key, value, exclude=None):
        self.kv[key] = value


class RefShapedPolicy(sb3_torch.ModularMlpPolicy):
    """The torch modules of oracle/sb3_torch.py behind the attribute names ModularPolicy's methods use; the
    methods themselves are the reference's."""
    _get_latent = ModularPolicy._get_latent
    _get_action_dist_from_latent = ModularPolicy._get_action_dist_from_latent
    evaluate_actions = ModularPolicy.evaluate_actions
    get_action_logits_from_obs = ModularPolicy.get_action_logits_from_obs
    forward = ModularPolicy.forward
    sde_features_extractor = None
    nomain = False
    log_std = None

    def extract_features(self, obs):  # SB3 preprocess_obs + FlattenExtractor
        return self.features(obs.long() if self.nvec is not None else obs)

    def mlp_extractor(self, features):
        return self.policy_net(features), self.value_net_body(features)

    @property
    def partner_mlp_extractor(self):
        return [lambda x, p=p: (self.partner_policy_net[p](x), self.partner_value_body[p](x))
                for p in range(self.num_partners)]


def run_reference_modular_train(kw, n_partners, Ms, BS, E, seed, coef, head_scale):
    import gym
    pol = RefShapedPolicy(nvec=kw["nvec"], heads=kw["heads"], num_partners=n_partners, seed=seed)
    nh, nslot = len(kw["heads"]), len(kw["nvec"])
    pol.action_dist = MultiCategoricalDistribution(kw["heads"]) if nh > 1 else CategoricalDistribution()
    with th.no_grad():  # the 0.01 head gain makes main and composed marginals equal (regulariser = 0): sharpen them
        pol.action_net.weight.mul_(head_scale)
        for p in range(n_partners):
            pol.partner_action_net[p].weight.mul_(head_scale * (p + 1))
    p0 = pol.to_flat().copy()
    out = dict(p0=p0, hp=np.array([n_partners, BS, E], np.int64), coef=np.array([coef], np.float64),
               Ms=np.array(Ms, np.int64))
    bufs = []
    for p, M in enumerate(Ms):
        obs, act, _, adv, ret = make_batch(kw, M, seed=seed + 1 + p)
        with th.no_grad():
            values, logp, _ = pol.evaluate_actions(th.as_tensor(obs[:, :nslot].astype(np.float32)),
                                                   th.as_tensor(act[:, :nh].astype(np.int64)) if nh > 1
                                                   else th.as_tensor(act[:, 0].astype(np.int64)), partner_idx=p)
        old_logp = (logp.numpy() + 0.1 * np.random.RandomState(seed + p).randn(M)).astype(np.float32)
        perms = oupd.perm_feistel(M, E, seed=seed, stream=4 + p)
        bufs.append(Buffer(obs[:, :nslot].astype(np.float32), act[:, :nh] if nh > 1 else act[:, :1], old_logp, adv, ret,
                           values.numpy().reshape(-1).astype(np.float32), perms))
        for k, v in dict(obs=obs, act=act, old_logp=old_logp, adv=adv, ret=ret, perms=perms).items():
            out[f"b{p}_{k}"] = v
    algo = Data(policy=pol, n_epochs=E, batch_size=BS, clip_range=lambda _: 0.2, clip_range_vf=None,
                _current_progress_remaining=1.0, _update_learning_rate=lambda opt: None, use_sde=False,
                action_space=gym.spaces.MultiDiscrete(kw["heads"]) if nh > 1 else gym.spaces.Discrete(kw["heads"][0]),
                ent_coef=0.01, vf_coef=0.5, marginal_reg_coef=coef, target_kl=None, verbose=0, max_grad_norm=0.5,
                _n_updates=0, logger=Log(), rollout_buffer=bufs)
    pol.parameters = pol.ordered_parameters  # what clip_grad_norm_ walks
    modular_learn.ModularAlgorithm.train(algo)  # <- the reference's own code
    out.update(params=pol.to_flat().copy(), log_keys=np.array(sorted(algo.logger.kv)),
               log_vals=np.array([float(algo.logger.kv[k]) for k in sorted(algo.logger.kv)], np.float64))
    return out


# ---------------------------------------------------------------------------------------------------------
# ModularAlgorithm.collect_rollouts (learn.py:155-218), executed verbatim around the reference's own MultiAgentEnv
# (LiarEnv, two scripted partners) behind a stand-in for DummyVecEnv: rows, who the partner was, bootstrap arguments.
def collect_fixture():
    import gym
    from pantheonrl.common.agents import Agent
    from pantheonrl.envs.liargym.liar import LiarEnv

    class Script(Agent):
        def __init__(self, actions):
            self.actions, self.k = actions, 0

        def get_action(self, obs, record=True):
            a = self.actions[self.k % len(self.actions)]
            self.k += 1
            return a

        def update(self, reward, done):
            pass

    class RecordingLiar(LiarEnv):
        def __init__(self):
            super().__init__()
            self.resets = []

        def multi_reset(self, egofirst):
            o = super().multi_reset(egofirst)
            self.resets.append([int(egofirst)] + [int(x) for x in self.egohand] + [int(x) for x in self.althand])
            return o

    class OneEnvVec:
        num_envs = 1

        def __init__(self, env):
            self.env, self.envs = env, [env]

        def reset(self):
            return np.asarray(self.env.reset())[None]

        def step(self, actions):
            obs, rew, done, info = self.env.step(actions[0])
            if done:
                obs = self.env.reset()
            return np.asarray(obs)[None], np.array([rew], np.float32), np.array([done]), [info]

    class ScriptedPolicy:
        def __init__(self, actions):
            self.actions, self.k, self.partner_of_call = actions, 0, []

        def forward(self, obs_tensor, partner_idx):
            a = np.asarray(self.actions[self.k % len(self.actions)]).reshape(1, -1)
            k = self.k
            self.k += 1
            self.partner_of_call.append(int(partner_idx))
            return th.as_tensor(a), th.tensor([[0.25 * k]]), th.tensor([-0.5 * k])

    class RecBuffer:
        def __init__(self):
            self.rollouts = []

        def reset(self):
            self.rollouts.append(dict(rows=[], gae=None))

        def add(self, obs, actions, rewards, dones, values, log_probs):
            start = -1.0 if dones is None else float(np.asarray(dones).reshape(-1)[0])  # learn.py:181: None on the first row
            self.rollouts[-1]["rows"].append((np.asarray(obs).reshape(-1).copy(), np.asarray(actions).reshape(-1).copy(),
                                              float(rewards[0]), start, float(values.reshape(-1)[0]),
                                              float(log_probs.reshape(-1)[0])))

        def compute_returns_and_advantage(self, last_values, dones):
            self.rollouts[-1]["gae"] = (float(last_values.reshape(-1)[0]), float(np.asarray(dones).reshape(-1)[0]))

    rng = np.random.RandomState(31)
    ego_script = [np.array([rng.randint(6), c]) for c in (1, 3, 5, 7, 9, 11, 2, 11, 4)] + [np.array([6, 11])]
    alt_scripts = [[np.array([rng.randint(6), c]) for c in (2, 4, 6, 8, 10, 1, 11, 3)] + [np.array([6, 11])],
                   [np.array([rng.randint(6), c]) for c in (3, 5, 2, 9)] + [np.array([6, 11])]]
    np.random.seed(323)
    base = RecordingLiar()
    partners = [Script(a) for a in alt_scripts]
    for p in partners:
        base.add_partner_agent(p)
    venv = OneEnvVec(base)
    cb = Data(on_rollout_start=lambda: None, on_rollout_end=lambda: None, on_step=lambda: True)
    pol = ScriptedPolicy(ego_script)
    algo = Data(_last_obs=venv.reset(), _last_dones=None, use_sde=False, sde_sample_freq=-1, policy=pol, device="cpu",
                action_space=gym.spaces.MultiDiscrete([7, 12]), num_timesteps=0, _update_info_buffer=lambda infos: None)
    n_steps, n_iter = 8, 2
    bufs = [RecBuffer(), RecBuffer()]
    partner_calls = []
    for _ in range(n_iter):  # ModularAlgorithm.learn (learn.py:378-384): one rollout per partner, in order
        for partner_idx in range(2):
            venv.envs[0].set_partnerid(partner_idx)
            assert modular_learn.ModularAlgorithm.collect_rollouts(algo, venv, cb, bufs[partner_idx], n_steps, partner_idx)
            partner_calls.append([p.k for p in partners])
    out = {}
    for q, buf in enumerate(bufs):
        R = buf.rollouts
        out.update({f"p{q}_row_obs": np.array([[r[0] for r in x["rows"]] for x in R]),
                    f"p{q}_row_act": np.array([[r[1] for r in x["rows"]] for x in R]),
                    f"p{q}_row_rew": np.array([[r[2] for r in x["rows"]] for x in R]),
                    f"p{q}_row_start": np.array([[r[3] for r in x["rows"]] for x in R]),
                    f"p{q}_row_value": np.array([[r[4] for r in x["rows"]] for x in R]),
                    f"p{q}_row_logp": np.array([[r[5] for r in x["rows"]] for x in R]),
                    f"p{q}_gae_last_value": np.array([x["gae"][0] for x in R]),
                    f"p{q}_gae_dones": np.array([x["gae"][1] for x in R])})
    out.update(num_timesteps=np.array(algo.num_timesteps), policy_calls=np.array(pol.k),
               partner_of_call=np.array(pol.partner_of_call), partner_calls=np.array(partner_calls),
               resets=np.array(base.resets), ego_script=np.array(ego_script), alt_script0=np.array(alt_scripts[0]),
               alt_script1=np.array(alt_scripts[1]), hp=np.array([n_steps, n_iter]))
    np.savez_compressed(os.path.join(HERE, "collect_rollouts_modular.npz"), **out)
    print("collect:", out["p0_row_obs"].shape, "first-row starts", out["p0_row_start"][:, 0], out["p1_row_start"][:, 0],
          "partner calls", out["partner_calls"].tolist())


def main():
    out = {}
    for name, kw, n_partners, Ms, BS, E, seed, coef, hs in (
            ("rps2", oracle.RPS_SPACE, 2, (200, 150), 64, 2, 5, 0.5, 100.0),
            ("liar3", oracle.LIAR_SPACE, 3, (300, 300, 260), 128, 2, 9, 1.0, 30.0),
            ("liar1", oracle.LIAR_SPACE, 1, (400,), 256, 2, 11, 0.0, 1.0)):
        for k, v in run_reference_modular_train(kw, n_partners, Ms, BS, E, seed, coef, hs).items():
            out[f"{name}_{k}"] = v
        print(name, dict(zip(out[name + "_log_keys"], np.round(out[name + "_log_vals"], 5))))
    np.savez_compressed(os.path.join(HERE, "modular.npz"), **out)


if __name__ == "__main__":
    main()
    collect_fixture()
