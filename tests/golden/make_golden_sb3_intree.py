"""Pin the ARITHMETIC half of the oracle on the reference's own in-tree code
(authoring container only:  python tests/golden/make_golden_sb3_intree.py).

stable-baselines3 itself is not installable here, but the reference vendors near-copies of the
two SB3 routines on the hot path, and those can be EXECUTED:

  * PPO.train  -> pantheonrl/algos/adap/adap_learn.py:229-347 (`ADAP.train`): the reference's copy of
    SB3 1.7.0 `PPO.train` plus one extra loss term, `context_loss_coeff * context_loss`.  The
    method is run UNBOUND on a duck-typed `self` (hyper-parameters, a rollout buffer that yields
    the minibatches of a stored permutation, a logger that captures `record`) with the policy
    of oracle/sb3_torch.py, `context_loss_coeff = 0` and `get_context_kl_loss` returning zero.
  * GAE        -> overcookedgym/human_aware_rl/baselines/baselines/ppo2/runner.py:152-164: the
    bootstrap loop, exec'd from the file's own source lines on float32 arrays.

  * BC         -> pantheonrl/algos/bc.py:270-357 (`BC._calculate_loss`, `BC.train`), run unbound the same way.

Writes sb3_intree.npz: inputs, the parameters / Adam moments / logged scalars after the
reference's train(), and the advantages / returns of its GAE loop.  tests/test_oracle_sb3_intree.py
replays them through oracle/sb3_torch.py, oracle/sb3_numpy.py and the C oracle.
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
REF = ref_shim.REF


def stub(name, **attrs):
    m = sys.modules.get(name) or types.ModuleType(name)
    sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def explained_variance(y_pred, y_true):
    var_y = np.var(y_true)
    return np.nan if var_y == 0 else 1 - np.var(y_true - y_pred) / var_y


# names adap_learn.py imports at module scope (none of them is executed by train())
stub("stable_baselines3.common.type_aliases", GymEnv=object, MaybeCallback=object, Schedule=object)
stub("stable_baselines3.common.utils", explained_variance=explained_variance, get_schedule_fn=lambda v: (lambda _: v),
     obs_as_tensor=None)
stub("stable_baselines3.common.vec_env", VecEnv=object)
stub("stable_baselines3.common.callbacks", BaseCallback=object)
stub("stable_baselines3.common.buffers", RolloutBuffer=object)
stub("pantheonrl.algos.adap.util", SAMPLERS={}, get_context_kl_loss=lambda algo, policy, data: th.zeros(()))
stub("pantheonrl.algos.adap.policies", AdapPolicy=object)
from pantheonrl.algos.adap import adap_learn  # noqa: E402  (the reference's file, verbatim)

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
        self.t = dict(observations=obs, actions=th.as_tensor(act).float(), old_log_prob=th.as_tensor(old_logp),
                      advantages=th.as_tensor(adv), returns=th.as_tensor(ret), old_values=th.as_tensor(old_values))
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


def run_reference_train(kw, M, BS, E, seed):
    import gym
    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=seed)
    p0 = pol.to_flat().copy()
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=seed + 1)
    space = oracle.make_space(**kw)
    ev = oracle.policy_forward(space, p0, obs, action_in=act)
    old_logp = (ev["logp"] + 0.1 * np.random.RandomState(seed).randn(M)).astype(np.float32)
    perms = oupd.perm_feistel(M, E, seed=seed, stream=4)
    nslot, nh = len(kw["nvec"]), len(kw["heads"])
    algo = Data(policy=pol, n_epochs=E, batch_size=BS, clip_range=lambda _: 0.2, clip_range_vf=None,
                _current_progress_remaining=1.0, _update_learning_rate=lambda opt: None, use_sde=False,
                action_space=gym.spaces.MultiDiscrete(kw["heads"]) if nh > 1 else gym.spaces.Discrete(kw["heads"][0]),
                ent_coef=0.01, vf_coef=0.5, context_loss_coeff=0.0, target_kl=None, verbose=0, max_grad_norm=0.5,
                _n_updates=0, logger=Log(),
                rollout_buffer=Buffer(obs[:, :nslot], act[:, :nh] if nh > 1 else act[:, :1], old_logp, adv, ret,
                                      ev["value"].astype(np.float32), perms))
    pol.parameters = pol.ordered_parameters  # what clip_grad_norm_ walks
    adap_learn.ADAP.train(algo)  # <- the reference's own code
    return dict(p0=p0, obs=obs, act=act, old_logp=old_logp, adv=adv, ret=ret, perms=perms,
                params=pol.to_flat().copy(), n_updates=np.array(algo._n_updates),
                log_keys=np.array(sorted(algo.logger.kv)),
                log_vals=np.array([float(algo.logger.kv[k]) for k in sorted(algo.logger.kv)], np.float64),
                hp=np.array([M, BS, E], np.int64))


def run_reference_gae(T, N, p_done, seed):
    src = open(os.path.join(REF, "overcookedgym/human_aware_rl/baselines/baselines/ppo2/runner.py")).read().splitlines()
    block = src[151:164]  # lines 152-164: "mb_returns = np.zeros_like" ... "mb_returns = mb_advs + mb_values"
    assert block[0].strip().startswith("mb_returns = np.zeros_like") and block[-1].strip() == "mb_returns = mb_advs + mb_values"
    code = "\n".join(ln[8:] for ln in block)  # drop the method body's indentation
    rng = np.random.RandomState(seed)
    ns = dict(np=np,
              mb_rewards=rng.randint(-1, 2, (T, N)).astype(np.float32), mb_values=rng.randn(T, N).astype(np.float32),
              mb_dones=rng.rand(T, N) < p_done, last_values=rng.randn(N).astype(np.float32),
              self=Data(nsteps=T, gamma=0.99, lam=0.95, dones=rng.rand(N) < p_done))
    exec(code, ns)
    return dict(rewards=ns["mb_rewards"], values=ns["mb_values"], episode_starts=ns["mb_dones"].astype(np.float32),
                last_values=ns["last_values"], dones=ns["self"].dones.astype(np.float32), advantages=ns["mb_advs"],
                returns=ns["mb_returns"])


def run_reference_bc(kw, M, E, seed, ent_weight, l2_weight):
    """BC.train / BC._calculate_loss of pantheonrl/algos/bc.py:270-357, executed unbound on a duck-typed
    self: batches of BC.DEFAULT_BATCH_SIZE cut from one stored permutation per epoch, torch's default Adam."""
    stub("stable_baselines3.common.policies", BasePolicy=object)
    import gym
    gym.Space = gym.spaces.Space  # only named in bc.py's annotations
    lg = Data(record=lambda *a, **k: None, dump=lambda *a, **k: None)
    u = sys.modules["stable_baselines3.common.utils"]
    u.configure_logger = lambda verbose=0, *a, **k: lg
    u.get_device = lambda d="auto": th.device("cpu")
    sb3c = sys.modules["stable_baselines3.common"]
    sb3c.policies, sb3c.utils = sys.modules["stable_baselines3.common.policies"], u
    from pantheonrl.algos import bc as ref_bc  # the reference's file, verbatim

    pol = sb3_torch.MlpPolicy(nvec=kw["nvec"], heads=kw["heads"], seed=seed)
    p0 = pol.to_flat().copy()
    obs, act, _, _, _ = make_batch(kw, M, seed=seed + 1)
    nslot, nh = len(kw["nvec"]), len(kw["heads"])
    perms = oupd.perm_feistel(M, E, seed=seed, stream=6)
    BS = ref_bc.BC.DEFAULT_BATCH_SIZE

    class Loader:
        def __init__(self):
            self.epoch = 0

        def __iter__(self):
            perm = np.asarray(perms[self.epoch])
            self.epoch += 1
            for s0 in range(0, M, BS):
                idx = perm[s0:s0 + BS]
                yield {"obs": obs[idx][:, :nslot].astype(np.int64), "acts": act[idx][:, :nh].astype(np.int64)}

    pol.parameters = pol.ordered_parameters
    losses = []
    me = Data(expert_data_loader=Loader(), policy=pol, device=th.device("cpu"), ent_weight=ent_weight,
              l2_weight=l2_weight, optimizer=th.optim.Adam(pol.ordered_parameters()))

    def calc(o, a):
        loss, stats = ref_bc.BC._calculate_loss(me, o, a)
        losses.append([stats["neglogp"], stats["entropy"], stats["prob_true_act"], stats["loss"], stats["l2_norm"]])
        return loss, stats
    me._calculate_loss = calc
    ref_bc.BC.train(me, n_epochs=E)  # <- the reference's own training loop
    return dict(p0=p0, obs=obs, act=act, perms=perms, params=pol.to_flat().copy(), stats=np.array(losses, np.float64),
                hp=np.array([M, BS, E], np.int64), w=np.array([ent_weight, l2_weight], np.float64))


def main():
    out = {}
    for name, kw, M, E, seed, ew, l2 in (("rps", oracle.RPS_SPACE, 100, 2, 3, 1e-3, 0.0),
                                         ("liar", oracle.LIAR_SPACE, 150, 2, 5, 1e-3, 0.01)):
        for k, v in run_reference_bc(kw, M, E, seed, ew, l2).items():
            out[f"bc_{name}_{k}"] = v
    for name, kw, M, BS, E, seed in (("rps", oracle.RPS_SPACE, 300, 64, 3, 7), ("liar", oracle.LIAR_SPACE, 700, 256, 2, 11)):
        for k, v in run_reference_train(kw, M, BS, E, seed).items():
            out[f"train_{name}_{k}"] = v
    for name, (T, N, p) in (("a", (128, 37, 0.25)), ("b", (2048, 1, 1.0)), ("c", (400, 5, 0.0025))):
        for k, v in run_reference_gae(T, N, p, seed=T + N).items():
            out[f"gae_{name}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "sb3_intree.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items() if "log" in k or "hp" in k})
    print("train_liar log:", dict(zip(out["train_liar_log_keys"], np.round(out["train_liar_log_vals"], 5))))


if __name__ == "__main__":
    main()
