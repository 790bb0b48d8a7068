"""Pin the ADAP half of the oracle on the reference's OWN code (authoring container only:
python tests/golden/make_golden_adap.py).

Executed verbatim from /root/reference (stable-baselines3 itself is not installable here, so SB3's base classes
are stand-ins; every line that computes something on this path is the reference's):

  * ADAP.train                 pantheonrl/algos/adap/adap_learn.py:229-347, unbound on a duck-typed `self`
  * get_context_kl_loss        pantheonrl/algos/adap/util.py:97-131 (+ kl_divergence :16-39, SAMPLERS :42-94)
  * AdapPolicy._get_latent / evaluate_actions / set_context / get_context
                               pantheonrl/algos/adap/policies.py:65-131, bound to the policy of oracle/sb3_torch.py
                               (whose towers take features ++ context like AdapPolicy._build_mlp_extractor :71-84)

SB3 stand-ins (restated from the 1.7.0 release): CategoricalDistribution / MultiCategoricalDistribution
(proba_distribution, log_prob, entropy over torch Categorical), preprocess_obs (one-hot), RolloutBuffer.get.

The random draws of the context loss (th.randperm for the states, SAMPLERS[...] for the contexts, both from
torch's global generator) are recorded while the reference runs, and stored with the result: the oracle and
the CUDA kernel take them as inputs.  Writes adap.npz.
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


def explained_variance(y_pred, y_true):
    var_y = np.var(y_true)
    return np.nan if var_y == 0 else 1 - np.var(y_true - y_pred) / var_y


class Distribution:
    pass


class CategoricalDistribution(Distribution):
    """SB3 1.7.0 distributions.CategoricalDistribution (the methods this path calls)."""

    def __init__(self, action_dim):
        self.action_dim = action_dim

    def proba_distribution(self, action_logits):
        self.distribution = th.distributions.Categorical(logits=action_logits)
        return self

    def log_prob(self, actions):
        return self.distribution.log_prob(actions)

    def entropy(self):
        return self.distribution.entropy()


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


dist_mod = stub("stable_baselines3.common.distributions", Distribution=Distribution,
                CategoricalDistribution=CategoricalDistribution,
                MultiCategoricalDistribution=MultiCategoricalDistribution)
sys.modules["stable_baselines3.common"].distributions = dist_mod
stub("stable_baselines3.common.type_aliases", GymEnv=object, MaybeCallback=object, Schedule=object)
stub("stable_baselines3.common.utils", explained_variance=explained_variance, get_schedule_fn=lambda v: (lambda _: v),
     obs_as_tensor=None, get_device=lambda d="auto": th.device("cpu"))
stub("stable_baselines3.common.vec_env", VecEnv=object)
stub("stable_baselines3.common.callbacks", BaseCallback=object)
stub("stable_baselines3.common.buffers", RolloutBuffer=object, RolloutBufferSamples=object)
stub("stable_baselines3.common.torch_layers", BaseFeaturesExtractor=object, FlattenExtractor=object,
     MlpExtractor=th.nn.Module)  # MultModel(MlpExtractor) calls nn.Module.__init__ itself (policies.py:134-147)
from pantheonrl.algos.adap import adap_learn, util as adap_util  # noqa: E402  (the reference's files, verbatim)
from pantheonrl.algos.adap.policies import AdapPolicy, MultModel  # noqa: E402

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
        self.values, self.returns = np.asarray(old_values), np.asarray(ret)

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


class RefShapedPolicy(sb3_torch.AdapMlpPolicy):
    """The torch modules of oracle/sb3_torch.py behind the attribute names the reference's AdapPolicy methods
    use; the methods themselves are the reference's."""
    _get_latent = AdapPolicy._get_latent
    evaluate_actions = AdapPolicy.evaluate_actions
    set_context = AdapPolicy.set_context
    get_context = AdapPolicy.get_context
    sde_features_extractor = None

    def extract_features(self, obs):  # SB3 preprocess_obs + FlattenExtractor
        return self.features(obs.long() if self.nvec is not None else obs)

    def mlp_extractor(self, features):
        return self.policy_net(features), self.value_net_body(features)

    def _get_action_dist_from_latent(self, latent_pi, latent_sde=None):  # SB3 ActorCriticPolicy, categorical cases
        return self.action_dist.proba_distribution(action_logits=self.action_net(latent_pi))


class RefShapedMultPolicy(sb3_torch.AdapMultPolicy):
    """AdapPolicyMult: AdapPolicy's own methods around the reference's own MultModel (policies.py:134-267), whose
    layers are replaced by the modules of oracle/sb3_torch.AdapMultPolicy (same shapes, built in the same order)."""
    _get_latent = AdapPolicy._get_latent
    evaluate_actions = AdapPolicy.evaluate_actions
    set_context = AdapPolicy.set_context
    get_context = AdapPolicy.get_context
    sde_features_extractor = None

    def __init__(self, **kw):
        super().__init__(**kw)
        mm = MultModel(self.F, [dict(pi=[64, 64], vf=[64, 64])], th.nn.Tanh, "cpu", self.context_size)
        for name in ("agent_branch_1", "agent_scaling", "agent_branch_2", "value_branch_1", "value_scaling",
                     "value_branch_2"):
            assert [tuple(p.shape) for p in getattr(mm, name).parameters()] == \
                [tuple(p.shape) for p in getattr(self, name).parameters()], name
            setattr(mm, name, getattr(self, name))
        object.__setattr__(self, "_mm", mm)  # not a registered submodule: the parameters are already ours

    def extract_features(self, obs):
        return self.features(obs.long() if self.nvec is not None else obs)

    def mlp_extractor(self, features):
        return self._mm(features)  # MultModel.forward: splits the context off, policies() / values()

    def _get_action_dist_from_latent(self, latent_pi, latent_sde=None):
        return self.action_dist.proba_distribution(action_logits=self.action_net(latent_pi))


def run_reference_adap_train(kw, M, BS, E, seed, K, S, coeff, sampler, head_scale=1.0, mult=False):
    import gym
    C = 3
    pol = (RefShapedMultPolicy if mult else RefShapedPolicy)(nvec=kw["nvec"], heads=kw["heads"], context_size=C, seed=seed)
    nh = len(kw["heads"])
    pol.action_dist = MultiCategoricalDistribution(kw["heads"]) if nh > 1 else CategoricalDistribution(kw["heads"][0])
    with th.no_grad():  # SB3's 0.01 head gain makes every context's distribution uniform (KL = 0): sharpen it
        pol.action_net.weight.mul_(head_scale)
    p0 = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=seed + 1)
    rs = np.random.RandomState(seed + 2)
    # the context stored with a sample stays the same over an episode: piecewise-constant unit vectors
    ctx = np.repeat(adap_util.get_L2_sphere(C, (M + 6) // 7, torch=True).numpy(), 7, axis=0)[:M].astype(np.float32)
    space = oracle.make_space(**kw)
    nslot0, nh0 = len(kw["nvec"]), len(kw["heads"])
    with th.no_grad():  # old log-probs / values from the policy itself (any numbers near the truth would do)
        v_, lp_, _ = pol.evaluate_actions(th.as_tensor(np.concatenate([obs[:, :nslot0].astype(np.float32), ctx], axis=1)),
                                          th.as_tensor(act[:, :nh0].astype(np.int64)) if nh0 > 1
                                          else th.as_tensor(act[:, 0].astype(np.int64)))
    ev = {"logp": lp_.numpy(), "value": v_.numpy().reshape(-1)}
    old_logp = (ev["logp"] + 0.1 * rs.randn(M)).astype(np.float32)
    perms = oupd.perm_feistel(M, E, seed=seed, stream=4)
    nslot = len(kw["nvec"])
    full_obs = np.concatenate([obs[:, :nslot].astype(np.float32), ctx], axis=1)
    algo = Data(policy=pol, n_epochs=E, batch_size=BS, clip_range=lambda _: 0.2, clip_range_vf=None,
                _current_progress_remaining=1.0, _update_learning_rate=lambda opt: None, use_sde=False,
                action_space=gym.spaces.MultiDiscrete(kw["heads"]) if nh > 1 else gym.spaces.Discrete(kw["heads"][0]),
                ent_coef=0.01, vf_coef=0.5, context_loss_coeff=coeff, target_kl=None, verbose=0, max_grad_norm=0.5,
                _n_updates=0, logger=Log(), context_size=C, num_context_samples=K, num_state_samples=S,
                context_sampler=sampler,
                rollout_buffer=Buffer(full_obs, act[:, :nh] if nh > 1 else act[:, :1], old_logp, adv, ret,
                                      ev["value"].astype(np.float32), perms))
    pol.parameters = pol.ordered_parameters  # what clip_grad_norm_ walks
    pol.set_context(th.as_tensor(ctx[:1]))
    # record the draws of get_context_kl_loss while it runs
    sidx, draws = [], []
    real_randperm, real_sampler = th.randperm, adap_util.SAMPLERS[sampler]

    def randperm(n, *a, **k):
        r = real_randperm(n, *a, **k)
        row = np.full(S, -1, np.int64)
        row[:min(S, n)] = r[:S].numpy()
        sidx.append(row)
        draws.append([])
        return r

    def sample(ctx_size, num, torch=False):
        c = real_sampler(ctx_size=ctx_size, num=num, torch=torch)
        draws[-1].append(np.asarray(c, np.float32).reshape(-1))
        return c
    th.manual_seed(seed + 3)
    th.randperm, adap_util.SAMPLERS[sampler] = randperm, sample
    try:
        adap_learn.ADAP.train(algo)  # <- the reference's own code, context loss included
    finally:
        th.randperm, adap_util.SAMPLERS[sampler] = real_randperm, real_sampler
    return dict(p0=p0, obs=obs, ctx=ctx, act=act, old_logp=old_logp, adv=adv, ret=ret, perms=perms,
                sidx=np.array(sidx, np.int32), draws=np.array(draws, np.float32), params=pol.to_flat().copy(),
                log_keys=np.array(sorted(algo.logger.kv)),
                log_vals=np.array([float(algo.logger.kv[k]) for k in sorted(algo.logger.kv)], np.float64),
                hp=np.array([M, BS, E, K, S], np.int64), coeff=np.array([coeff], np.float64))


def main():
    out = {}
    for name, kw, M, BS, E, seed, K, S, coeff, sampler, hs, mult in (
            ("rps", oracle.RPS_SPACE, 280, 64, 3, 7, 5, 32, 0.1, "l2", 200.0, False),      # last minibatch: 24 < S states
            ("liar", oracle.LIAR_SPACE, 700, 256, 2, 11, 5, 32, 0.5, "l2", 1.0, False),     # two context tiles per minibatch
            ("liar_k3", oracle.LIAR_SPACE, 300, 100, 2, 13, 3, 20, 1.0, "unit_square", 100.0, False),
            ("mult_rps", oracle.RPS_SPACE, 280, 64, 2, 17, 5, 32, 0.5, "l2", 200.0, True),  # AdapPolicyMult (ADAP_MULT)
            ("mult_liar", oracle.LIAR_SPACE, 500, 250, 2, 19, 4, 32, 1.0, "l2", 100.0, True)):
        for k, v in run_reference_adap_train(kw, M, BS, E, seed, K, S, coeff, sampler, hs, mult).items():
            out[f"{name}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "adap.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items() if "hp" in k or "sidx" in k or "draws" in k})
    for n in ("rps", "liar", "liar_k3", "mult_rps", "mult_liar"):
        print(n, "log:", dict(zip(out[n + "_log_keys"], np.round(out[n + "_log_vals"], 5))))


if __name__ == "__main__":
    main()
