"""GPU tests of the host side that keeps the reference's plugin surface:
MultiAgentEnv / TurnBasedEnv / SimultaneousEnv routing replayed against event
traces recorded from the reference's own classes (tests/golden), the
OnPolicyAgent lazy-train protocol, PPO.learn at n_envs = 1 and the hand-over to
the device engine at n_envs > 1, and a multi-iteration bit-exact cross-check of
the engine against the oracle pieces."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import rollout as orc
from oracle import update as oupd
from pantheonrl_b200 import _lib, ops, update as dupd
from pantheonrl_b200.common.agents import Agent, OnPolicyAgent, StaticPolicyAgent
from pantheonrl_b200.common.multiagentenv import PlayerException
from pantheonrl_b200.engine import PPOConfig, VecTrainer
from pantheonrl_b200.envs import LiarEnv, RPSEnv, make
from pantheonrl_b200.ppo import PPO

pytestmark = pytest.mark.gpu


class Scripted(Agent):
    def __init__(self, pid, actions, log):
        self.pid, self.actions, self.log, self.k = pid, actions, log, 0

    def get_action(self, obs, record=True):
        a = self.actions[self.k]
        self.k += 1
        self.log.append(("act", self.pid, np.asarray(obs.obs).reshape(-1).copy(), np.asarray(a).reshape(-1).copy()))
        return a

    def update(self, reward, done):
        self.log.append(("upd", self.pid, float(reward), bool(done)))


class ScriptedLiar(LiarEnv):
    """LiarEnv whose dice / coin come from a recorded list instead of the device RNG."""

    def __init__(self, resets):
        super().__init__()
        self.resets, self.r = resets, 0

    def draw_ego_first(self):
        return bool(self.resets[self.r][0])

    def multi_reset(self, egofirst):
        info = self.resets[self.r]
        self.r += 1
        st = np.zeros((1, 32), np.uint8)
        st[0, :12] = info[1:13]
        self.state = torch.from_numpy(st).cuda()
        hand = info[1:7] if egofirst else info[7:13]
        return np.concatenate([hand, np.tile([6, 0], 12)]).astype(np.int64)


def _replay(env, g, n_partners):
    log = []
    pid_of = g["ev_pid"][g["ev_kind"] == 0]
    acts = g["ev_act"][g["ev_kind"] == 0]
    for p in range(n_partners):
        a = [x if x.size > 1 else int(x[0]) for x in acts[pid_of == p]]
        env.add_partner_agent(Scripted(p, a, log))
    obs = env.reset()
    T = g["ego_act"].shape[0]
    for t in range(T):
        assert np.array_equal(np.asarray(obs).reshape(-1), g["ego_obs"][t]), t
        a = g["ego_act"][t]
        o2, r, d, info = env.step(a if a.size > 1 else int(a[0]))
        assert r == g["ego_rew"][t] and int(d) == g["ego_done"][t], t
        assert info["_partnerid"][0] == g["ego_pid"][t]
        obs = env.reset() if d else o2
    assert np.array_equal(np.asarray(obs).reshape(-1), g["final_obs"])
    assert len(log) == len(g["ev_kind"])
    for i, e in enumerate(log):
        assert (0 if e[0] == "act" else 1) == g["ev_kind"][i] and e[1] == g["ev_pid"][i], i
        if e[0] == "act":
            assert np.array_equal(e[2], g["ev_obs"][i]) and np.array_equal(e[3], g["ev_act"][i]), i
        else:
            assert e[2] == g["ev_rew"][i] and int(e[3]) == g["ev_done"][i], i


@pytest.mark.parametrize("fname,n_partners", [("routing_liar.npz", 1), ("routing_liar_rr3.npz", 3)])
def test_turnbased_routing_matches_reference_trace(ctx, golden_dir, fname, n_partners):
    g = dict(np.load(os.path.join(golden_dir, fname)))
    _replay(ScriptedLiar(g["reset_info"]), g, n_partners)


def test_simultaneous_routing_matches_reference_trace(ctx, golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "routing_rps.npz")))
    _replay(RPSEnv(), g, 1)


def test_player_exceptions_match_the_reference_contract(ctx):
    env = RPSEnv()
    with pytest.raises(PlayerException):
        env.add_partner_agent(Scripted(0, [], []), player_num=0)  # the ego's seat
    with pytest.raises(PlayerException):
        env.set_resample_policy("bogus")
    assert make("LiarsDice-v0").observation_space.nvec.sum() == 270


def test_liar_env_device_rng_is_reproducible(ctx):
    a, b = LiarEnv(seed=3), LiarEnv(seed=3)
    for _ in range(5):
        assert a.draw_ego_first() == b.draw_ego_first()
        oa, ob = a.multi_reset(True), b.multi_reset(True)
        assert np.array_equal(oa, ob) and sum(a.hands[0]) == 6 and sum(a.hands[1]) == 6
    st, ef, obs = oracle.liar_reset(1, seed=3, tick=0)
    c = LiarEnv(seed=3)
    assert c.draw_ego_first() == bool(ef[0])
    c.multi_reset(bool(ef[0]))
    assert np.array_equal(c.state.cpu().numpy()[0, :12], st[0, :12])


def test_onpolicy_agent_lazy_train_protocol(ctx):
    """agents.py:123-184: the partner trains inside the get_action that follows its
    n_steps-th recorded action, bootstrapping from the last stored value."""
    env = RPSEnv()
    model = PPO("MlpPolicy", env, n_steps=8, batch_size=4, n_epochs=2, seed=10)
    agent = OnPolicyAgent(model)
    from pantheonrl_b200.common.observation import Observation
    p0 = model.policy.params.clone()
    rew = [1, -1, 0, 1, 1, -1, 0, 0]
    for t in range(8):
        agent.get_action(Observation(np.array([0])))
        agent.update(0, False)       # first-move hand-off
        agent.update(rew[t], True)
    buf = model.rollout_buffer
    assert buf.pos == 8 and model._n_updates == 0 and torch.equal(model.policy.params, p0)
    assert np.array_equal(buf.h["rewards"], np.array(rew, np.float32))
    assert np.array_equal(buf.h["episode_starts"], np.ones(8, np.float32))
    vals, starts, lps = buf.h["values"].copy(), buf.h["episode_starts"].copy(), buf.h["logp"].copy()
    agent.get_action(Observation(np.array([0])))          # 9th call: GAE + train + reset, then record row 0
    assert model._n_updates == 2 and buf.pos == 1 and not torch.equal(model.policy.params, p0)
    adv, ret = oracle.gae(np.array(rew, np.float32)[:, None], vals[:, None], starts[:, None], vals[-1:], np.ones(1))
    assert np.array_equal(buf.d["advantages"].cpu().numpy(), adv)
    agent.update(5.0, False)
    model.rollout_buffer.reset()
    agent.update(7.0, True)  # cursor 0: dropped like the reference drops it (agents.py:198)
    assert model.rollout_buffer.h["rewards"][0] == 5.0


def test_ppo_learn_single_env_and_static_partner(ctx, tmp_path):
    env = LiarEnv(seed=1)
    partner = OnPolicyAgent(PPO("MlpPolicy", env, n_steps=32, batch_size=16, n_epochs=2, seed=10))
    env.add_partner_agent(partner)
    ego = PPO("MlpPolicy", env, n_steps=32, batch_size=16, n_epochs=2, seed=10)
    assert torch.equal(ego.policy.params, partner.model.policy.params)  # same seed -> same init (trainer.py:111,198)
    ego.learn(total_timesteps=96)
    assert ego.num_timesteps == 96 and ego._n_updates == 6
    assert partner.model._n_updates >= 2 and partner.num_timesteps >= 40
    st = ego.last_stats.cpu().numpy()
    assert np.all(np.isfinite(st)) and bool(torch.isfinite(ego.policy.params).all())
    # save / load round trip and a FIXED partner built from the loaded policy (trainer.py:140-162)
    path = str(tmp_path / "ego.pt")
    ego.save(path)
    again = PPO.load(path, env)
    assert torch.equal(again.policy.params, ego.policy.params)
    env2 = LiarEnv(seed=2)
    env2.add_partner_agent(StaticPolicyAgent(again.policy))
    o = env2.reset()
    for _ in range(20):
        o, r, d, _ = env2.step(np.array([0, 11]))
        if d:
            o = env2.reset()


def _scalars(run_dir):
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(str(run_dir))
    acc.Reload()
    return {t: [(e.step, e.value) for e in acc.Scalars(t)] for t in acc.Tags()["scalars"]}


SB3_TRAIN_TAGS = {"train/entropy_loss", "train/policy_gradient_loss", "train/value_loss", "train/approx_kl",
                  "train/clip_fraction", "train/loss", "train/explained_variance", "train/clip_range"}


def test_tensorboard_scalars_and_sb3_zip(ctx, tmp_path):
    """SURVEY.md 8f-2 / 8f-3 through the facade: the TensorBoard tags / steps SB3 writes for the ego
    (learn) and for an OnPolicyAgent partner (agents.py:132-153), and an SB3-shaped archive that
    PPO.load(path) reopens without an env (trainer.py:149)."""
    tb = tmp_path / "tb"
    env = RPSEnv()
    partner = OnPolicyAgent(PPO("MlpPolicy", env, n_steps=16, batch_size=8, n_epochs=2, seed=10), log_interval=1, tensorboard_log=str(tb),
                            tb_log_name="partner")
    env.add_partner_agent(partner)
    ego = PPO("MlpPolicy", env, n_steps=16, batch_size=8, n_epochs=2, seed=10, tensorboard_log=str(tb))
    ego.learn(total_timesteps=48, tb_log_name="ego")
    ego.logger.close()
    partner.model.logger.close()
    sc = _scalars(tb / "ego_1")
    assert SB3_TRAIN_TAGS | {"rollout/ep_rew_mean", "rollout/ep_len_mean", "time/fps"} <= set(sc)
    assert [s for s, _ in sc["rollout/ep_len_mean"]] == [16, 32, 48]  # dumped before each train(), step = timesteps
    assert all(v == 1.0 for _, v in sc["rollout/ep_len_mean"])        # RPS: every step ends an episode
    assert [s for s, _ in sc["train/loss"]] == [32, 48]               # train() scalars ride on the next dump
    stats = ego.last_stats.cpu().numpy()
    ps = _scalars(tb / "partner_1")
    assert {"rollout/ep_rew_mean", "rollout/ep_len_mean"} <= set(ps) and [s for s, _ in ps["rollout/ep_rew_mean"]][:2] == [16, 32]
    assert np.isfinite(stats).all()
    # SB3-shaped archive, reopened the way trainer.py does for LOAD / FIXED partners
    import zipfile
    path = ego.save(str(tmp_path / "ego_model"))
    assert path.endswith("ego_model.zip")
    assert {"data", "policy.pth", "policy.optimizer.pth"} <= set(zipfile.ZipFile(path).namelist())
    again = PPO.load(str(tmp_path / "ego_model"))
    assert torch.equal(again.policy.params, ego.policy.params) and torch.equal(again.adam_m, ego.adam_m)
    assert torch.equal(again.adam_v, ego.adam_v)
    assert (again.adam_step, again._n_updates, again.num_timesteps) == (ego.adam_step, ego._n_updates, 48)
    assert again.n_steps == 16 and again.observation_space.n == 1 and again.action_space.n == 3
    env2 = RPSEnv()
    env2.add_partner_agent(StaticPolicyAgent(again.policy))
    again.set_env(env2)
    again.learn(total_timesteps=16)  # training continues from the loaded optimizer state
    assert again._n_updates == ego._n_updates + 2
    # device loop (n_envs > 1): same tags, one dump per iteration
    env3 = LiarEnv()
    p3 = OnPolicyAgent(PPO("MlpPolicy", env3, n_steps=8, n_epochs=1, seed=1, n_minibatches=2), log_interval=1, tensorboard_log=str(tb), tb_log_name="p3")
    env3.add_partner_agent(p3)
    e3 = PPO("MlpPolicy", env3, n_steps=8, n_epochs=1, seed=1, n_envs=256, n_minibatches=2, tensorboard_log=str(tb))
    e3.learn(total_timesteps=256 * 8 * 2, tb_log_name="vec")
    e3.logger.close()
    p3.model.logger.close()
    v = _scalars(tb / "vec_1")
    assert SB3_TRAIN_TAGS <= set(v) and [s for s, _ in v["rollout/ep_rew_mean"]] == [2048, 4096]
    assert all(1.0 <= x <= 13.0 for _, x in v["rollout/ep_len_mean"])
    assert SB3_TRAIN_TAGS <= set(_scalars(tb / "p3_1"))


def test_wrappers_around_device_games(ctx, tmp_path):
    """--framestack / --record of trainer.py:92-100 on the device-backed games: the ego trains on
    stacked observations while every step is recorded and written in the reference's .npy layout."""
    from pantheonrl_b200.common import trajsaver, wrappers
    env = RPSEnv()
    env.add_partner_agent(OnPolicyAgent(PPO("MlpPolicy", env, n_steps=16, batch_size=8, n_epochs=1, seed=10)))
    env = wrappers.frame_wrap(env, 3)
    env = rec = wrappers.recorder_wrap(env)
    assert env.observation_space.nvec.tolist() == [1, 1, 1]
    ego = PPO("MlpPolicy", env, n_steps=16, batch_size=8, n_epochs=1, seed=10)
    ego.learn(total_timesteps=32)
    tr = rec.get_transitions()
    assert tr.egoobs.shape == (32, 3) and tr.flags.tolist() == [wrappers.DONE] * 32  # RPS: one step per game
    tr.write_transition(tmp_path / "rps.npy")
    back = trajsaver.SimultaneousTransitions.read_transition(tmp_path / "rps.npy", env.observation_space,
                                                             env.action_space)
    assert np.array_equal(back.egoacts.reshape(-1), tr.egoacts) and set(np.unique(tr.altacts)) <= {0, 1, 2}
    # turn based, recorded
    lenv = LiarEnv(seed=3)
    lenv.add_partner_agent(StaticPolicyAgent(PPO("MlpPolicy", lenv, seed=1).policy))
    lrec = wrappers.recorder_wrap(lenv)
    PPO("MlpPolicy", lrec, n_steps=16, batch_size=8, n_epochs=1, seed=2).learn(total_timesteps=16)
    lt = lrec.get_transitions()
    assert len(lt.get_ego_transitions()) == 16 and lt.obs.shape[1] == 30
    assert len(lt.get_alt_transitions()) == int((lt.flags % 2 == 1).sum()) > 0
    # trainer.py --framestack 3 on Liar's Dice: 90 observation slots (rows of 96 bytes), ego and partner both
    # learn on stacked frames (trainer.py:97-99 wraps env and altenv)
    senv = wrappers.frame_wrap(LiarEnv(seed=4), 3)
    assert len(senv.observation_space.nvec) == 90
    spartner = OnPolicyAgent(PPO("MlpPolicy", senv, n_steps=24, batch_size=8, n_epochs=2, seed=10))
    senv.add_partner_agent(spartner)
    sego = PPO("MlpPolicy", senv, n_steps=32, batch_size=16, n_epochs=2, seed=10)
    p0 = sego.policy.params.clone()
    sego.learn(total_timesteps=96)
    assert sego._n_updates == 6 and spartner.model._n_updates >= 2
    assert bool(torch.isfinite(sego.policy.params).all()) and not torch.equal(sego.policy.params, p0)
    assert sego.rollout_buffer.h["obs"].shape == (32, 96)
    with pytest.raises(_lib.PthError):
        PPO("MlpPolicy", wrappers.frame_wrap(LiarEnv(), 4), seed=1)  # 120 slots: more than PTH_MAX_OBS_SLOTS


def test_behaviour_cloning_facade(ctx, tmp_path):
    """pantheonrl/algos/bc.py surface: clone a scripted Liar's Dice partner from recorded transitions
    (recorder_wrap -> get_alt_transitions -> BC.train), save / reconstruct the policy (trainer.py:152)."""
    from pantheonrl_b200.bc import BC, BCShell, reconstruct_policy
    from pantheonrl_b200.common import wrappers
    from pantheonrl_b200.envs.liar import LiarDefaultAgent
    lenv = LiarEnv(seed=5)
    lenv.add_partner_agent(LiarDefaultAgent())
    rec = wrappers.recorder_wrap(lenv)
    o = rec.reset()
    rng = np.random.RandomState(0)
    for _ in range(400):  # a random ego keeps the games going; the scripted partner is the expert
        o, r, d, _ = rec.step(np.array([rng.randint(6), rng.randint(12)]))
        if d:
            o = rec.reset()
    expert = rec.get_transitions().get_alt_transitions()
    assert len(expert) > 100 and expert.obs.shape[1] == 30
    bc = BC(lenv.observation_space, lenv.action_space, expert_data=expert, seed=1)
    p0 = bc.policy.params.clone()
    bc.train(n_epochs=30)
    per_epoch = -(-len(expert) // BC.DEFAULT_BATCH_SIZE)
    st = bc.last_stats.cpu().numpy()
    assert st.shape == (30 * per_epoch, 8) and bc.adam_step == 30 * per_epoch
    first, last = st[:per_epoch, 0].mean(), st[-per_epoch:, 0].mean()
    assert last < 0.6 * first, (first, last)                      # neglogp of the expert's actions drops
    assert st[-per_epoch:, 3].mean() > 2 * st[:per_epoch, 3].mean()  # prob_true_act rises
    row = bc.stats_row(st[-1])
    assert row["loss"] == pytest.approx(row["neglogp"] - 1e-3 * row["entropy"])
    assert not torch.equal(bc.policy.params, p0)
    bc.train(n_batches=per_epoch + 2)  # one full epoch + the first two batches of another
    assert bc.last_stats.shape[0] == per_epoch + 2 and bc.batches_done == 31 * per_epoch + 2
    with pytest.raises(ValueError):
        bc.train(n_epochs=1, n_batches=1)
    # the value tower is untouched by BC (vf_coef = 0): Adam sees exact zeros there
    lay = bc.policy.state_dict()
    assert torch.equal(lay["value_net.weight"], BC(lenv.observation_space, lenv.action_space, seed=1).policy.state_dict()["value_net.weight"])
    bc.save_policy(str(tmp_path / "bc_policy.pt"))
    pol2 = reconstruct_policy(str(tmp_path / "bc_policy.pt"))
    assert torch.equal(pol2.params, bc.policy.params)
    env2 = LiarEnv(seed=6)
    env2.add_partner_agent(StaticPolicyAgent(BCShell(pol2).policy))  # gen_fixed for a BC partner
    o = env2.reset()
    for _ in range(10):
        o, r, d, _ = env2.step(np.array([1, 11]))
        if d:
            o = env2.reset()


def test_ppo_learn_on_device_matches_engine(ctx):
    N, T = 256, 16
    env = LiarEnv()
    partner = OnPolicyAgent(PPO("MlpPolicy", env, n_steps=T, n_epochs=2, seed=10, n_minibatches=4))
    env.add_partner_agent(partner)
    ego = PPO("MlpPolicy", env, n_steps=T, n_epochs=2, seed=10, n_envs=N, n_minibatches=4)
    ego.learn(total_timesteps=N * T * 2)
    cfg = PPOConfig(n_steps=T, n_epochs=2, n_minibatches=4)
    tr = VecTrainer("liar", N, cfg, seed=10, partner="ppo")
    tr.learn(N * T * 2)
    assert torch.equal(ego.policy.params, tr.ego.params)
    assert torch.equal(partner.model.policy.params, tr.alt.params)
    # self-play: partner = StaticPolicyAgent(ego.policy)
    env = RPSEnv()
    ego = PPO("MlpPolicy", env, n_steps=8, n_epochs=1, seed=3, n_envs=512, n_minibatches=2)
    env.add_partner_agent(StaticPolicyAgent(ego.policy))
    ego.learn(total_timesteps=512 * 8)
    assert ego._trainer.alt is None and ego.num_timesteps == 512 * 8


def test_engine_three_iterations_bit_exact_vs_oracle(ctx):
    """The north star's trace claim across training: rollout -> GAE -> PPO.train ->
    rollout with the UPDATED weights, all buffers and parameters equal to the oracle."""
    N, T, E, NMB, seed = 192, 12, 2, 3, 7
    cfg = PPOConfig(n_steps=T, n_epochs=E, n_minibatches=NMB)
    tr = VecTrainer("liar", N, cfg, seed=seed, partner="ppo")
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    sp = tr.space
    pe = tr.ego.params.cpu().numpy().copy()
    pa = tr.alt.params.cpu().numpy().copy()
    me, ve, ma, va = (np.zeros_like(pe) for _ in range(4))
    step_e = step_a = upd_e = upd_a = 0
    carry = None
    o_alt = orc.new_buffer(orc.alt_capacity("liar", T), N, True)  # one partner buffer: open rows are carried in it
    for it in range(3):
        tr.iteration()
        torch.cuda.synchronize()
        o_ego, o_alt, carry = orc.rollout("liar", osp, pe, pa, N=N, T=T, seed=seed, tick0=it * T,
                                          first_rollout=it == 0, carry=carry, alt=o_alt)
        for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
            assert np.array_equal(getattr(tr.ego_buf, k).cpu().numpy(), o_ego[k]), (it, k)
        assert np.array_equal(tr.alt_buf.count.cpu().numpy(), o_alt["count"])
        adv, ret = oracle.gae(o_ego["rewards"], o_ego["values"], o_ego["episode_starts"],
                              carry["ego_last_value"], carry["ego_last_done"])
        aadv, aret = oracle.gae_ragged(o_alt["rewards"], o_alt["values"], o_alt["episode_starts"],
                                       o_alt["count"], carry["alt_boot_done"])
        # ego update
        idx = oupd.index_build(None, T, N)
        M = idx.size
        bs = -(-M // NMB)
        perm = oupd.perm_feistel(M, E, seed, _lib.STREAM_SHUFFLE_EGO, epoch0=upd_e)
        G = tr.last_grids[0] or dupd.update_grid(sp, M, bs)  # the grid is part of the reduction contract
        oupd.ppo_update(osp, pe, me, ve, step_e, o_ego["obs"], o_ego["actions"], o_ego["logp"], adv, ret, perm, bs, G,
                        index=idx)
        step_e += E * (-(-M // bs))
        upd_e += E
        # partner update
        aidx = oupd.index_build(o_alt["count"], orc.alt_capacity("liar", T), N)
        M = aidx.size
        bs = -(-M // NMB)
        perm = oupd.perm_feistel(M, E, seed, _lib.STREAM_SHUFFLE_ALT, epoch0=upd_a)
        G = tr.last_grids[1] or dupd.update_grid(sp, M, bs)
        oupd.ppo_update(osp, pa, ma, va, step_a, o_alt["obs"], o_alt["actions"], o_alt["logp"], aadv, aret, perm, bs,
                        G, index=aidx)
        step_a += E * (-(-M // bs))
        upd_a += E
        assert np.array_equal(tr.ego.params.cpu().numpy(), pe), f"ego params differ after iteration {it}"
        assert np.array_equal(tr.alt.params.cpu().numpy(), pa), f"partner params differ after iteration {it}"
    st = tr.train_stats()
    assert np.isfinite(st["train/loss"]) and tr.episode_stats()["ego_steps"] == 3 * N * T


def test_load_then_learn_on_device_continues_counters(ctx, tmp_path):
    """PPO.load followed by learn() with n_envs > 1: the device trainer takes over the loaded Adam step
    count (bias correction), update counter (shuffle keys) and starts its env / sampling streams at fresh
    ticks, for the ego and for the partner."""
    N, T = 128, 8
    env = LiarEnv()
    partner = OnPolicyAgent(PPO("MlpPolicy", env, n_steps=T, n_epochs=2, seed=10, n_minibatches=4))
    env.add_partner_agent(partner)
    ego = PPO("MlpPolicy", env, n_steps=T, n_epochs=2, seed=10, n_envs=N, n_minibatches=4)
    ego.learn(total_timesteps=N * T)
    first_perm = ego._trainer.ego_perm.clone()
    ego.save(str(tmp_path / "ego"))
    partner.model.save(str(tmp_path / "alt"))
    env2 = LiarEnv()
    partner2 = OnPolicyAgent(PPO.load(str(tmp_path / "alt"), env2))
    env2.add_partner_agent(partner2)
    ego2 = PPO.load(str(tmp_path / "ego"), env2)
    assert (ego2.adam_step, ego2._n_updates, ego2.n_envs) == (ego.adam_step, 2, N) and ego.adam_step == 8
    assert partner2.model.adam_step == partner.model.adam_step > 0
    ego2.learn(total_timesteps=N * T)
    tr = ego2._trainer
    assert tr.tick_base == 2 * T and (tr.ego.adam_step, tr.ego.n_updates) == (16, 4)
    assert (ego2.adam_step, ego2._n_updates, ego2.num_timesteps) == (16, 4, N * T)
    assert partner2.model._n_updates == 4 and partner2.model.adam_step == tr.alt.adam_step > partner.model.adam_step
    assert not torch.equal(tr.ego_perm, first_perm)  # epochs 2, 3 of the shuffle stream, not 0, 1 again
    # the same continuation driven by hand: a fresh VecTrainer given the saved state and counters
    cfg = PPOConfig(n_steps=T, n_epochs=2, n_minibatches=4)
    ref = VecTrainer("liar", N, cfg, seed=10, partner="ppo")
    ref.ego.params.copy_(ego.policy.params)
    ref.ego.adam_m.copy_(ego.adam_m)
    ref.ego.adam_v.copy_(ego.adam_v)
    ref.alt.params.copy_(partner.model.policy.params)
    ref.alt.adam_m.copy_(partner.model.adam_m)
    ref.alt.adam_v.copy_(partner.model.adam_v)
    ref.ego.adam_step, ref.ego.n_updates = ego.adam_step, 2
    ref.alt.adam_step, ref.alt.n_updates = partner.model.adam_step, 2
    ref.tick_base = 2 * T
    ref.iteration()
    assert torch.equal(ref.ego.params, ego2.policy.params) and torch.equal(ref.alt.params, partner2.model.policy.params)


def test_recorded_transitions_from_the_device_buffers(ctx):
    """VecTrainer.recorded_transitions (device -> host copies + vec_record) gives, for any env of the
    engine, the transitions cut from the ORACLE's buffers of the same rollout — which
    tests/test_vec_record_cpu.py pins byte for byte on the reference's own recorder."""
    from pantheonrl_b200 import vec_record as vr
    N, T, seed = 96, 20, 5
    cfg = PPOConfig(n_steps=T, n_epochs=1, n_minibatches=2)
    tr = VecTrainer("liar", N, cfg, seed=seed, partner="ppo")
    pe, pa = tr.ego.params.cpu().numpy().copy(), tr.alt.params.cpu().numpy().copy()
    tr.collect()
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    o_ego, o_alt, carry = orc.rollout("liar", osp, pe, pa, N=N, T=T, seed=seed)
    o_alt = dict(o_alt, count=o_alt["count"] + ((carry["flags"] >> 2) & 1))  # + the row left open at the end
    for env in (0, 41, N - 1):
        got = tr.recorded_transitions(env)
        want = vr.turn_based_transitions(o_ego, o_alt, env, carry["ego_last_done"][env])
        assert np.array_equal(got.obs, want.obs) and np.array_equal(got.acts, want.acts)
        assert np.array_equal(got.flags, want.flags) and len(got.get_ego_transitions()) == T
    tr.collect()
    with pytest.raises(_lib.PthError):
        tr.recorded_transitions(0)  # a later turn-based rollout may start inside an episode
    # simultaneous game: any rollout
    tr = VecTrainer("rps", N, cfg, seed=seed, partner="ppo")
    pe, pa = tr.ego.params.cpu().numpy().copy(), tr.alt.params.cpu().numpy().copy()
    tr.collect()
    rsp = oracle.make_space(**oracle.RPS_SPACE)
    o_ego, o_alt, carry = orc.rollout("rps", rsp, pe, pa, N=N, T=T, seed=seed)
    got = tr.recorded_transitions(7)
    want = vr.simultaneous_transitions(o_ego, o_alt, 7, carry["ego_last_done"][7], obs_len=1)
    for name in ("egoobs", "egoacts", "altobs", "altacts", "flags"):
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
    with pytest.raises(_lib.PthError):  # a static partner stores no rows — also for Liar's Dice
        VecTrainer("liar", 8, cfg, seed=1, partner="selfplay").recorded_transitions(0)


def test_partner_set_three_partners_on_one_gpu_bit_exact_vs_oracle(ctx):
    """Several partner learners per GPU (the reference's partner set, trainer.py:216-228): envs cut into
    one group per partner, one ego update over all groups, P partner updates — every buffer and every
    parameter equal to the oracle pieces run group by group with the same global env ids."""
    from pantheonrl_b200.dist_util import global_env_major_index
    from pantheonrl_b200.partner_set import PartnerSetTrainer
    P, Ng, T, E, NMB, seed = 3, 64, 10, 2, 3, 11
    N = P * Ng
    cfg = PPOConfig(n_steps=T, n_epochs=E, n_minibatches=NMB)
    tr = PartnerSetTrainer("liar", N, cfg, partners_per_gpu=P, seed=seed)
    osp, sp = oracle.make_space(**oracle.LIAR_SPACE), tr.space
    pe = tr.ego.params.cpu().numpy().copy()
    pas = [ln.learner.params.cpu().numpy().copy() for ln in tr.lanes]
    me, ve = np.zeros_like(pe), np.zeros_like(pe)
    mas, vas = [np.zeros_like(pe) for _ in range(P)], [np.zeros_like(pe) for _ in range(P)]
    step_e = upd_e = 0
    step_a, carries = [0] * P, [None] * P
    o_alts = [orc.new_buffer(orc.alt_capacity("liar", T), Ng, True) for _ in range(P)]
    idx = global_env_major_index(P, T, Ng).numpy()
    for it in range(2):
        tr.iteration()
        torch.cuda.synchronize()
        cat = {k: [] for k in ("obs", "actions", "logp", "adv", "ret")}
        parts = []
        for p in range(P):
            o_ego, o_alts[p], carries[p] = orc.rollout("liar", osp, pe, pas[p], N=Ng, T=T, seed=seed, tick0=it * T,
                                                       env0=p * Ng, first_rollout=it == 0, carry=carries[p],
                                                       alt=o_alts[p])
            ln = tr.lanes[p]
            for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
                assert np.array_equal(getattr(ln.ego_view, k).cpu().numpy(), o_ego[k]), (it, p, k)
            assert np.array_equal(ln.buf.count.cpu().numpy(), o_alts[p]["count"])
            adv, ret = oracle.gae(o_ego["rewards"], o_ego["values"], o_ego["episode_starts"],
                                  carries[p]["ego_last_value"], carries[p]["ego_last_done"])
            for k, v in (("obs", o_ego["obs"]), ("actions", o_ego["actions"]), ("logp", o_ego["logp"]),
                         ("adv", adv), ("ret", ret)):
                cat[k].append(v.reshape(T * Ng, -1) if v.ndim == 3 else v.reshape(T * Ng))
            parts.append((o_alts[p], carries[p]))
        cat = {k: np.concatenate(v) for k, v in cat.items()}
        M = idx.size
        bs = -(-M // NMB)
        perm = oupd.perm_feistel(M, E, seed, _lib.STREAM_SHUFFLE_EGO, epoch0=upd_e)
        oupd.ppo_update(osp, pe, me, ve, step_e, cat["obs"], cat["actions"], cat["logp"], cat["adv"], cat["ret"], perm,
                        bs, tr.last_grids[0], index=idx)
        step_e += E * (-(-M // bs))
        assert np.array_equal(tr.ego.params.cpu().numpy(), pe), f"ego params differ after iteration {it}"
        for p, (o_alt, carry) in enumerate(parts):
            aadv, aret = oracle.gae_ragged(o_alt["rewards"], o_alt["values"], o_alt["episode_starts"], o_alt["count"],
                                           carry["alt_boot_done"])
            aidx = oupd.index_build(o_alt["count"], orc.alt_capacity("liar", T), Ng)
            Ma = aidx.size
            bsa = -(-Ma // NMB)
            aperm = oupd.perm_feistel(Ma, E, seed, tr.lanes[p].shuffle_stream, epoch0=upd_e)
            oupd.ppo_update(osp, pas[p], mas[p], vas[p], step_a[p], o_alt["obs"], o_alt["actions"], o_alt["logp"],
                            aadv, aret, aperm, bsa, tr.last_grids[1 + p], index=aidx)
            step_a[p] += E * (-(-Ma // bsa))
            assert np.array_equal(tr.lanes[p].learner.params.cpu().numpy(), pas[p]), (it, p)
        upd_e += E
    assert not np.array_equal(pas[0], pas[1]) and tr.episode_stats()["ego_steps"] == 2 * N * T
    # the facade: three OnPolicyAgent partners added to one env -> the same trainer
    env = LiarEnv()
    agents = [OnPolicyAgent(PPO("MlpPolicy", env, n_steps=T, n_epochs=E, seed=seed, n_minibatches=NMB)) for _ in range(P)]
    for a in agents:
        env.add_partner_agent(a)
    ego = PPO("MlpPolicy", env, n_steps=T, n_epochs=E, seed=seed, n_envs=N, n_minibatches=NMB)
    ego.learn(total_timesteps=2 * N * T)
    assert np.array_equal(ego.policy.params.cpu().numpy(), pe)
    for p in range(P):
        assert np.array_equal(agents[p].model.policy.params.cpu().numpy(), pas[p])
