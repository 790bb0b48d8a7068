"""smoke(): one tiny invocation of the whole hot path on cuda:0 — rollout of both
agents, GAE, one PPO.train per learner — checked bit for bit against the CPU
oracle (the oracle is only the checker here)."""
import numpy as np
import torch


def run():
    import oracle
    from oracle import rollout as orc
    from oracle import update as oupd
    from . import _lib, ops, rollout as dev, update as dupd
    from .policy import init_flat

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    N, T, seed = 256, 8, 10
    sp = dev.space_for("liar")
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    flat = init_flat(sp, seed)
    pe = torch.from_numpy(flat).cuda()
    pa = pe.clone()
    ego, alt, carry = dev.Buffer(T, N, False, "cuda"), dev.Buffer(dev.alt_capacity("liar", T), N, True, "cuda"), dev.Carry(N, "cuda")
    dev.run_rollout("liar", sp, pe, pa, ego, alt, carry, T, seed, 0, first_rollout=True)
    adv, ret = ops.gae(ego.rewards, ego.values, ego.episode_starts, carry.ego_last_value, carry.ego_last_done)
    index, _ = dupd.index_build(None, T, N)
    M, bs, grid = T * N, 512, 3
    perm = dupd.perm_feistel(M, 2, seed, _lib.STREAM_SHUFFLE_EGO)
    ws = dupd.UpdateWorkspace(sp, M, bs)
    m, v = torch.zeros_like(pe), torch.zeros_like(pe)
    p_new = pe.clone()
    stats = dupd.ppo_update(sp, p_new, m, v, 0, ego.obs, ego.actions, ego.logp, adv, ret, perm, bs, ws,
                            index=index, grid_ctas=grid)
    torch.cuda.synchronize()

    # ---- the same on the CPU oracle
    o_ego, o_alt, o_carry = orc.rollout("liar", osp, flat, flat, N=N, T=T, seed=seed)
    for k in ("obs", "actions", "rewards", "values", "logp", "episode_starts"):
        assert np.array_equal(getattr(ego, k).cpu().numpy(), o_ego[k]), f"rollout {k} differs from the oracle"
    assert np.array_equal(alt.count.cpu().numpy(), o_alt["count"])
    o_adv, o_ret = oracle.gae(o_ego["rewards"], o_ego["values"], o_ego["episode_starts"],
                              o_carry["ego_last_value"], o_carry["ego_last_done"])
    assert np.array_equal(adv.cpu().numpy(), o_adv) and np.array_equal(ret.cpu().numpy(), o_ret)
    o_index = oupd.index_build(None, T, N)
    o_perm = oupd.perm_feistel(M, 2, seed, _lib.STREAM_SHUFFLE_EGO)
    assert np.array_equal(perm.cpu().numpy(), o_perm) and np.array_equal(index.cpu().numpy(), o_index)
    op, om, ov = flat.copy(), np.zeros_like(flat), np.zeros_like(flat)
    o_stats, _ = oupd.ppo_update(osp, op, om, ov, 0, o_ego["obs"], o_ego["actions"], o_ego["logp"], o_adv,
                                 o_ret, o_perm, bs, grid, index=o_index)
    assert np.array_equal(p_new.cpu().numpy(), op), "updated parameters differ from the oracle"
    assert np.array_equal(stats.cpu().numpy(), o_stats)
    print(f"smoke ok: liar N={N} T={T}: rollout + GAE + PPO.train bit-exact vs oracle; "
          f"loss={float(stats[-1, 5]):.5f} launches={_lib.LAUNCHES}")
