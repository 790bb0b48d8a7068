"""GPU parity for the Overcooked path (BASELINE configs[3]):
  * the device gridworld (pth_env_overcooked_*) against the reference-generated golden traces and,
    at N >> 1, against the CPU oracle with random joint actions on every layout;
  * the rollout megakernel's Overcooked variant, the Box forward and the Box PPO update,
    bit for bit against the oracle; two whole engine iterations against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import overcooked as ooc
from oracle import rollout as orc
from oracle import update as oupd
from pantheonrl_b200 import _lib, ops, rollout as dev, update as dupd
from pantheonrl_b200.engine import PPOConfig, VecTrainer
from pantheonrl_b200.envs import overcooked as oc
from test_oracle_cpu import rand_params

pytestmark = pytest.mark.gpu

OSPACE = dict(box_dim=62, heads=[6])


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _replay_device(layout, ego_idx, actions_player_order, horizon=400):
    """Replays [S][2] joint actions (player order) through pth_env_overcooked_step, N = 1,
    resetting after done.  Returns obs [S+1][2][62] in PLAYER order, rewards, dones, states."""
    L = oc.build_layout(layout, ego_idx, horizon)
    d_layout = oc.layout_to_device(L)
    state, obs = oc.env_reset(d_layout, 1)
    S = len(actions_player_order)
    out_obs, rew, done, states = [], [], [], []
    for i in range(S + 1):
        o = obs[0, :, :62].cpu().numpy()
        out_obs.append(o if ego_idx == 0 else o[::-1])
        states.append(state[0].cpu().numpy().copy())
        if i == S:
            break
        a0, a1 = actions_player_order[i]
        ea, aa = (a0, a1) if ego_idx == 0 else (a1, a0)
        obs, r, d = oc.env_step(d_layout, state, torch.tensor([ea], dtype=torch.uint8, device="cuda"),
                                torch.tensor([aa], dtype=torch.uint8, device="cuda"))
        rew.append(float(r.item()))
        done.append(int(d.item()))
        if done[-1]:
            state, obs = oc.env_reset(d_layout, 1)
    return np.array(out_obs), np.array(rew), np.array(done), np.array(states)


def test_device_env_reproduces_reference_featurization_pickle(ctx, golden_dir):
    """The reference's own golden vector (state_featurization.pickle) through the CUDA env."""
    g = _load(golden_dir, "oc_ref_featurization.npz")
    for ep in (0, 3):
        obs, rew, done, _ = _replay_device("simple", 0, g["actions"][ep])
        assert np.array_equal(obs[:400], g["feats"][ep].astype(np.float32))
        # OvercookedMultiEnv adds its shaping (3/3/5) to the sparse reward: strip the shaped part
        _, sparse, shaped, _, _ = ooc.replay(ooc.multienv_layout(golden_dir, "simple"), g["actions"][ep])
        assert np.array_equal(rew, (sparse + shaped).astype(np.float64))
        assert np.array_equal(sparse, g["sparse"][ep]) and done[-1] == 1


@pytest.mark.parametrize("layout,ego_idx", [("simple", 0), ("simple", 1), ("random1", 0), ("corridor", 1),
                                            ("scenario2_s", 0)])
def test_device_env_reproduces_reference_random_trace(ctx, golden_dir, layout, ego_idx):
    g = _load(golden_dir, f"oc_random_{layout}.npz")
    S = 700
    obs, rew, done, _ = _replay_device(layout, ego_idx, g["actions"][:S])
    assert np.array_equal(obs[:, 0], g["obs0"][:S + 1].astype(np.float32))
    assert np.array_equal(obs[:, 1], g["obs1"][:S + 1].astype(np.float32))
    assert np.array_equal(rew, g["rewards"][:S].astype(np.float64))
    assert np.array_equal(done, g["dones"][:S])


@pytest.mark.parametrize("layout", oc.LAYOUT_LIST)
def test_device_env_vectorised_vs_oracle_every_layout(ctx, golden_dir, layout):
    """N envs with independent random action streams vs the oracle, every trainer.py layout."""
    N, S, horizon = 96, 260, 120
    rng = np.random.RandomState(__import__('zlib').crc32(layout.encode()))
    acts = np.where(rng.rand(N, S, 2) < 0.35, 5, rng.randint(0, 5, (N, S, 2))).astype(np.uint8)
    L = oc.build_layout(layout, 0, horizon)
    d_layout = oc.layout_to_device(L)
    oL = ooc.multienv_layout(golden_dir, layout, horizon=horizon)
    want = [ooc.replay(oL, acts[n]) for n in range(N)]
    state, obs = oc.env_reset(d_layout, N)
    fresh_state, fresh_obs = state.clone(), obs.clone()
    d_acts = torch.from_numpy(acts).cuda()
    for i in range(S):
        w = np.stack([want[n][0][i] for n in range(N)])
        assert np.array_equal(obs[:, :, :62].cpu().numpy(), w), (layout, i)
        obs, r, d = oc.env_step(d_layout, state, d_acts[:, i, 0].contiguous(), d_acts[:, i, 1].contiguous())
        assert np.array_equal(r.cpu().numpy(), np.array([want[n][1][i] + want[n][2][i] for n in range(N)], np.float32))
        assert np.array_equal(d.cpu().numpy(), np.array([want[n][3][i] for n in range(N)], np.uint8))
        m = d.bool()
        state[m] = fresh_state[m]
        obs[m] = fresh_obs[m]
    if layout in ('simple', 'random1', 'scenario2_s'):
        assert float(sum(w[1].sum() + w[2].sum() for w in want)) > 0  # something got cooked / placed


# ------------------------------------------------------------------ rollout
def _gpu_rollout(pe, pa, N, T, seed, layout, ego_idx=0, horizon=400, tick0=0, env0=0, first=True, carry=None,
                 selfplay=False, records=True):
    sp = dev.space_for("overcooked")
    L = oc.build_layout(layout, ego_idx, horizon)
    d_layout = oc.layout_to_device(L)
    d_pe = torch.from_numpy(pe).cuda()
    d_pa = d_pe if selfplay else torch.from_numpy(pa).cuda()
    ego = dev.Buffer(T, N, False, "cuda", box=True)
    alt = dev.Buffer(T, N, True, "cuda", box=True)
    carry = carry or dev.Carry(N, "cuda", _lib.PTH_OC_STATE_BYTES)
    dev.run_rollout("overcooked", sp, d_pe, d_pa, ego, alt, carry, T, seed, tick0, env0=env0,
                    first_rollout=first, partner_records=records, d_layout=d_layout)
    torch.cuda.synchronize()
    return ego, alt, carry


def _oracle_state_as_device(o_state):
    """orc_oc_state rows -> the fields both encodings share: positions, orientations, held, t."""
    p = o_state[:, :14].reshape(-1, 2, 7)
    return p[:, :, 0], p[:, :, 1], p[:, :, 2], p[:, :, 3], o_state[:, 536:540].copy().view(np.int32)[:, 0]


def _compare(ego, alt, carry, o_ego, o_alt, o_carry, records=True):
    for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
        assert np.array_equal(getattr(ego, k).cpu().numpy(), o_ego[k]), f"ego {k}"
    for k in ("ego_last_start", "alt_last_done", "total_rew", "flags", "ego_last_value", "ego_last_done"):
        assert np.array_equal(getattr(carry, k).cpu().numpy(), o_carry[k]), f"carry {k}"
    assert np.array_equal(carry.ep_stats.cpu().numpy(), o_carry["ep_stats"])
    gs = carry.game_state.cpu().numpy()
    px, py, po, held, t = _oracle_state_as_device(o_carry["oc_state"])
    assert np.array_equal(gs[:, 16:18], px) and np.array_equal(gs[:, 18:20], py)
    assert np.array_equal(gs[:, 20:22], po) and np.array_equal(gs[:, 22:24], held)
    assert np.array_equal(gs[:, 32:34].copy().view(np.uint16)[:, 0], t)
    if records:
        cnt = alt.count.cpu().numpy()
        assert np.array_equal(cnt, o_alt["count"])
        for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
            assert np.array_equal(getattr(alt, k).cpu().numpy(), o_alt[k]), f"alt {k}"


@pytest.mark.parametrize("layout,ego_idx,N,T,horizon", [("simple", 0, 70, 50, 20), ("simple", 1, 33, 64, 400),
                                                        ("unident_s", 0, 64, 40, 15), ("corridor", 1, 40, 30, 400)])
def test_rollout_bit_exact_vs_oracle(ctx, golden_dir, layout, ego_idx, N, T, horizon):
    osp = oracle.make_space(**OSPACE)
    pe = rand_params(osp, seed=1, scale=0.3)
    pa = rand_params(osp, seed=2, scale=0.3)
    oL = ooc.multienv_layout(golden_dir, layout, horizon=horizon)
    o = orc.rollout("overcooked", osp, pe, pa, N=N, T=T, seed=10, tick0=5, env0=7, oc_layout=oL, oc_ego_idx=ego_idx)
    g = _gpu_rollout(pe, pa, N, T, 10, layout, ego_idx, horizon, tick0=5, env0=7)
    _compare(*g, *o)
    assert o[2]["ep_stats"][0] == (T // horizon) * N


def test_rollout_carry_and_selfplay(ctx, golden_dir):
    osp = oracle.make_space(**OSPACE)
    pe = rand_params(osp, seed=3, scale=0.4)
    pa = rand_params(osp, seed=4, scale=0.4)
    oL = ooc.multienv_layout(golden_dir, "simple", horizon=37)
    N, T = 100, 30
    o1 = orc.rollout("overcooked", osp, pe, pa, N=N, T=T, seed=2, oc_layout=oL)
    g1 = _gpu_rollout(pe, pa, N, T, 2, "simple", horizon=37)
    _compare(*g1, *o1)
    o2 = orc.rollout("overcooked", osp, pe, pa, N=N, T=T, seed=2, tick0=T, first_rollout=False, carry=o1[2], oc_layout=oL)
    g2 = _gpu_rollout(pe, pa, N, T, 2, "simple", horizon=37, tick0=T, first=False, carry=g1[2])
    _compare(*g2, *o2)
    # self-play: StaticPolicyAgent(ego.policy), nothing recorded for the partner
    o = orc.rollout("overcooked", osp, pe, pe, N=N, T=T, seed=1, partner_records=False, oc_layout=oL)
    g = _gpu_rollout(pe, pe, N, T, 1, "simple", horizon=37, selfplay=True, records=False)
    _compare(*g, *o, records=False)


def test_config4_size_rollout_properties(ctx, golden_dir):
    """BASELINE configs[3] size: 1024 envs, layout simple, horizon 400, T = 400."""
    osp = oracle.make_space(**OSPACE)
    pe = rand_params(osp, seed=8, scale=0.2)
    pa = rand_params(osp, seed=9, scale=0.2)
    N, T = 1024, 400
    ego, alt, carry = _gpu_rollout(pe, pa, N, T, 10, "simple")
    assert bool((ego.episode_starts[0] == 1).all()) and bool((ego.episode_starts[1:] == 0).all())
    assert bool((carry.ego_last_done == 1).all())                 # horizon reached on the last tick
    assert bool((alt.count == T).all())
    assert torch.equal(ego.rewards, alt.rewards)                   # one reward for both agents (overcooked.py:80)
    r = ego.rewards
    assert bool(((r >= 0) & (r == r.round())).all()) and float(r.sum()) > 0
    st = carry.ep_stats.cpu().numpy()
    assert st[0] == N and st[2] == N * T and st[3] == N * T
    assert bool((ego.obs[..., 62:] == 0).all()) and bool((ego.obs == ego.obs.round()).all())
    oL = ooc.multienv_layout(golden_dir, "simple")
    for n in (0, 517, 1023):  # sampled envs replayed on the oracle
        o = orc.rollout("overcooked", osp, pe, pa, N=1, T=T, seed=10, env0=n, oc_layout=oL)
        assert np.array_equal(ego.actions[:, n].cpu().numpy(), o[0]["actions"][:, 0])
        assert np.array_equal(ego.rewards[:, n].cpu().numpy(), o[0]["rewards"][:, 0])
        assert np.array_equal(ego.obs[:, n].cpu().numpy(), o[0]["obs"][:, 0])


# ------------------------------------------------------------------ Box forward / update
def _box_batch(M, seed):
    rng = np.random.RandomState(seed)
    obs = np.zeros((M, 64), np.float32)
    obs[:, :62] = rng.randint(-4, 5, (M, 62))
    act = np.zeros((M, 4), np.uint8)
    act[:, 0] = rng.randint(0, 6, M)
    adv = rng.randn(M).astype(np.float32)
    ret = rng.randn(M).astype(np.float32)
    return obs, act, adv, ret


def test_box_forward_bit_exact(ctx):
    osp = oracle.make_space(**OSPACE)
    sp = _lib.Space.box(62, [6])
    params = rand_params(osp, seed=11, scale=0.3)
    obs, act, _, _ = _box_batch(777, 1)
    want = oracle.policy_forward(osp, params, obs, seed=3, tick=9, idx0=5)
    got = ops.policy_forward(sp, torch.from_numpy(params).cuda(), torch.from_numpy(obs).cuda(), seed=3, tick=9, idx0=5)
    for k in ("action", "value", "logp", "entropy", "logits"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k


@pytest.mark.parametrize("M,BS,E,grid", [(300, 300, 1, 2), (700, 256, 3, 3), (2048, 512, 2, 4), (1000, 64, 2, 1)])
def test_box_update_bit_exact_vs_oracle(ctx, M, BS, E, grid):
    osp = oracle.make_space(**OSPACE)
    sp = _lib.Space.box(62, [6])
    params = rand_params(osp, seed=M, scale=0.3)
    obs, act, adv, ret = _box_batch(M, M)
    ev = oracle.policy_forward(osp, params, obs, action_in=act)
    old_logp = (ev["logp"] + 0.1 * np.random.RandomState(2).randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
    dp, dm, dv = d(params), d(np.zeros_like(params)), d(np.zeros_like(params))
    ws = dupd.UpdateWorkspace(sp, M, BS)
    gst = dupd.ppo_update(sp, dp, dm, dv, 0, d(obs), d(act), d(old_logp), d(adv), d(ret), d(perm), BS, ws,
                          grid_ctas=grid, ent_coef=0.01)
    torch.cuda.synchronize()
    op, om, ov = params.copy(), np.zeros_like(params), np.zeros_like(params)
    ost, _ = oupd.ppo_update(osp, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, ent_coef=0.01)
    assert np.array_equal(gst.cpu().numpy(), ost), np.abs(gst.cpu().numpy() - ost).max()
    assert np.array_equal(dm.cpu().numpy(), om) and np.array_equal(dv.cpu().numpy(), ov)
    assert np.array_equal(dp.cpu().numpy(), op), np.abs(dp.cpu().numpy() - op).max()


def test_box_update_matches_torch_autograd(ctx):
    """Independent check of the hand-written Box backward: torch autograd + torch Adam."""
    M, BS, E = 1500, 512, 2
    osp = oracle.make_space(**OSPACE)
    sp = _lib.Space.box(62, [6])
    from pantheonrl_b200 import policy
    flat = policy.init_flat(sp, 5)
    obs, act, adv, ret = _box_batch(M, 1)
    ev = oracle.policy_forward(osp, flat, obs, action_in=act)
    old_logp = (ev["logp"] + 0.05 * np.random.RandomState(3).randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=1, stream=5)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
    dp, dm, dv = d(flat), d(np.zeros_like(flat)), d(np.zeros_like(flat))
    ws = dupd.UpdateWorkspace(sp, M, BS)
    dupd.ppo_update(sp, dp, dm, dv, 0, d(obs), d(act), d(old_logp), d(adv), d(ret), d(perm), BS, ws)
    # torch restatement
    sd = policy.flat_to_state_dict(sp, flat)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(P.values()), lr=3e-4, eps=1e-5)
    x = torch.from_numpy(obs[:, :62]); a = torch.from_numpy(act[:, 0].astype(np.int64))
    t = lambda v: torch.from_numpy(v)  # noqa: E731
    lin = torch.nn.functional.linear
    for e in range(E):
        for m0 in range(0, M, BS):
            idx = torch.from_numpy(perm[e, m0:m0 + BS].astype(np.int64))
            xb = x[idx]
            hp = torch.tanh(lin(torch.tanh(lin(xb, P["mlp_extractor.policy_net.0.weight"], P["mlp_extractor.policy_net.0.bias"])),
                                P["mlp_extractor.policy_net.2.weight"], P["mlp_extractor.policy_net.2.bias"]))
            hv = torch.tanh(lin(torch.tanh(lin(xb, P["mlp_extractor.value_net.0.weight"], P["mlp_extractor.value_net.0.bias"])),
                                P["mlp_extractor.value_net.2.weight"], P["mlp_extractor.value_net.2.bias"]))
            dist = torch.distributions.Categorical(logits=lin(hp, P["action_net.weight"], P["action_net.bias"]))
            v = lin(hv, P["value_net.weight"], P["value_net.bias"]).flatten()
            advb = t(adv)[idx]
            advb = (advb - advb.mean()) / (advb.std() + 1e-8)
            ratio = torch.exp(dist.log_prob(a[idx]) - t(old_logp)[idx])
            pl = -torch.min(advb * ratio, advb * torch.clamp(ratio, 0.8, 1.2)).mean()
            loss = pl + 0.5 * torch.nn.functional.mse_loss(t(ret)[idx], v)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(P.values()), 0.5)
            opt.step()
    want = policy.state_dict_to_flat(sp, {k: v.detach() for k, v in P.items()})
    assert np.allclose(dp.cpu().numpy(), want, atol=5e-6, rtol=0)


def test_engine_two_iterations_bit_exact_vs_oracle(ctx, golden_dir):
    N, T, E, NMB, seed, horizon = 96, 40, 2, 3, 7, 25
    cfg = PPOConfig(n_steps=T, n_epochs=E, n_minibatches=NMB)
    tr = VecTrainer("overcooked", N, cfg, seed=seed, partner="ppo", layout="simple", horizon=horizon)
    osp = oracle.make_space(**OSPACE)
    oL = ooc.multienv_layout(golden_dir, "simple", horizon=horizon)
    sp = tr.space
    pe = tr.ego.params.cpu().numpy().copy()
    pa = tr.alt.params.cpu().numpy().copy()
    me, ve, ma, va = (np.zeros_like(pe) for _ in range(4))
    step = {"e": 0, "a": 0}
    upd = {"e": 0, "a": 0}
    carry = None
    for it in range(2):
        tr.iteration()
        torch.cuda.synchronize()
        o_ego, o_alt, carry = orc.rollout("overcooked", osp, pe, pa, N=N, T=T, seed=seed, tick0=it * T,
                                          first_rollout=it == 0, carry=carry, oc_layout=oL)
        for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
            assert np.array_equal(getattr(tr.ego_buf, k).cpu().numpy(), o_ego[k]), (it, k)
        adv, ret = oracle.gae(o_ego["rewards"], o_ego["values"], o_ego["episode_starts"],
                              carry["ego_last_value"], carry["ego_last_done"])
        aadv, aret = oracle.gae_ragged(o_alt["rewards"], o_alt["values"], o_alt["episode_starts"],
                                       o_alt["count"], carry["alt_boot_done"])
        for who, p, m, v, buf, a_, r_, stream, cnt in (
                ("e", pe, me, ve, o_ego, adv, ret, _lib.STREAM_SHUFFLE_EGO, None),
                ("a", pa, ma, va, o_alt, aadv, aret, _lib.STREAM_SHUFFLE_ALT, o_alt["count"])):
            idx = oupd.index_build(cnt, T, N)
            M = idx.size
            bs = -(-M // NMB)
            perm = oupd.perm_feistel(M, E, seed, stream, epoch0=upd[who])
            G = tr.last_grids[0 if who == "e" else 1] or dupd.update_grid(sp, M, bs)
            oupd.ppo_update(osp, p, m, v, step[who], buf["obs"], buf["actions"], buf["logp"], a_, r_, perm, bs, G,
                            index=idx)
            step[who] += E * (-(-M // bs))
            upd[who] += E
        assert np.array_equal(tr.ego.params.cpu().numpy(), pe), f"ego params differ after iteration {it}"
        assert np.array_equal(tr.alt.params.cpu().numpy(), pa), f"partner params differ after iteration {it}"
    assert tr.episode_stats()["episodes"] == N * (2 * T // horizon)


def test_facade_env_steps_on_device(ctx, golden_dir):
    """OvercookedMultiEnv (reference class name / signature) through MultiAgentEnv.step."""
    g = _load(golden_dir, "oc_routing_simple_e1.npz")
    from pantheonrl_b200.common.agents import Agent

    class Scripted(Agent):
        def __init__(self, acts):
            self.acts, self.k, self.seen, self.upd = acts, 0, [], []

        def get_action(self, obs, record=True):
            self.seen.append(np.array(obs.obs))
            self.k += 1
            return int(self.acts[self.k - 1])

        def update(self, reward, done):
            self.upd.append((reward, done))

    env = oc.OvercookedMultiEnv("simple", ego_agent_idx=1)
    env.layout.horizon = int(g["horizon"])
    env.d_layout = oc.layout_to_device(env.layout)
    alt_acts = g["ev_act"][g["ev_kind"] == 0]
    partner = Scripted(alt_acts)
    env.add_partner_agent(partner)
    obs = env.reset()
    T = 200
    for i in range(T):
        assert np.array_equal(np.asarray(obs), g["ego_obs"][i].astype(np.float64)), i
        o2, r, d, _ = env.step(int(g["ego_act"][i]))
        assert r == g["ego_rew"][i] and int(d) == g["ego_done"][i]
        obs = env.reset() if d else o2
    assert np.array_equal(np.array(partner.seen), g["ev_obs"][g["ev_kind"] == 0][:T].astype(np.float64))


def test_facade_ppo_learns_host_driven_and_on_device(ctx):
    """trainer.py-style drop-in: OvercookedMultiEnv + OnPolicyAgent(PPO) partner + ego PPO.learn at N = 1
    (one kernel call per decision, Box rows staged on the host), and the same objects with n_envs > 1
    handing the loop to the device engine."""
    from pantheonrl_b200.common.agents import OnPolicyAgent
    from pantheonrl_b200.ppo import PPO
    env = oc.OvercookedMultiEnv("simple")
    env.layout.horizon = 30
    env.d_layout = oc.layout_to_device(env.layout)
    partner = OnPolicyAgent(PPO("MlpPolicy", env, n_steps=40, batch_size=20, n_epochs=2, seed=10))
    env.add_partner_agent(partner)
    ego = PPO("MlpPolicy", env, n_steps=40, batch_size=20, n_epochs=2, seed=10)
    assert torch.equal(ego.policy.params, partner.model.policy.params)
    ego.learn(total_timesteps=80)
    assert ego.num_timesteps == 80 and ego._n_updates == 4 and partner.model._n_updates >= 2
    assert bool(torch.isfinite(ego.policy.params).all()) and np.all(np.isfinite(ego.last_stats.cpu().numpy()))
    # n_envs > 1: the device loop, equal to driving VecTrainer directly
    N, T = 64, 32
    env = oc.OvercookedMultiEnv("unident_s", ego_agent_idx=1)
    partner = OnPolicyAgent(PPO("MlpPolicy", env, n_steps=T, n_epochs=2, seed=10, n_minibatches=4))
    env.add_partner_agent(partner)
    ego = PPO("MlpPolicy", env, n_steps=T, n_epochs=2, seed=10, n_envs=N, n_minibatches=4)
    ego.learn(total_timesteps=N * T * 2)
    tr = VecTrainer("overcooked", N, PPOConfig(n_steps=T, n_epochs=2, n_minibatches=4), seed=10, partner="ppo",
                    layout="unident_s", ego_agent_idx=1)
    tr.learn(N * T * 2)
    assert torch.equal(ego.policy.params, tr.ego.params)
    assert torch.equal(partner.model.policy.params, tr.alt.params)
