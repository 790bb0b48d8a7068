"""Multi-GPU parity checks, one process per GPU (spawned by tests/test_gpu_multi.py with
torch.multiprocessing, or launched under torchrun by tools/multi_gpu_check.py):

  * pth_ppo_update with world > 1 — tiles sharded over ranks, the per-minibatch advantage
    statistics computed by one rank each, per-rank gradient sums exchanged through NVLink peer
    memory inside the persistent kernel — equals the CPU oracle's world-aware reduction bit for bit;
  * the engine at world > 1: ego replicas bit-identical on every rank, partners different learners,
    the gathered ego stream equal to every rank's own packed records, NCCL and fused pack + P2P
    exchange leading to the same parameters, sharded and replicated ego update both consistent.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def sharded_update_vs_oracle(rank, local, world, box=False):
    import oracle
    from oracle import update as oupd
    from pantheonrl_b200 import rollout as ro, update as dupd
    from pantheonrl_b200.policy import init_flat
    from test_oracle_update import make_batch
    kw = oracle.LIAR_SPACE
    sp, osp = ro.space_for("liar"), oracle.make_space(**kw)
    M, BS, E, grid = 1500, 400, 3, 2  # 4 minibatches x 3 epochs = 12 statistics ids over `world` ranks
    obs, act, old_logp, adv, ret = make_batch(kw, M, seed=3)
    params = init_flat(sp, 4)
    ev = oracle.policy_forward(osp, params, obs, action_in=act)
    old_logp = (ev["logp"] + 0.1 * np.random.RandomState(2).randn(M)).astype(np.float32)
    perm = oupd.perm_feistel(M, E, seed=10, stream=4)
    dev = f"cuda:{local}"
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    dp, dm, dv = d(params), torch.zeros(params.size, device=dev), torch.zeros(params.size, device=dev)
    peers = dupd.PeerExchange(sp, dist.group.WORLD, dev)
    ws = dupd.UpdateWorkspace(sp, M, BS, dev)
    for _ in range(3):  # three launches: epoch tags and buffer parities carry over
        st = dupd.ppo_update(sp, dp, dm, dv, 0, d(obs), d(act), d(old_logp), d(adv), d(ret), d(perm), BS, ws,
                             grid_ctas=grid, peers=peers, ent_coef=0.01)
    torch.cuda.synchronize()
    op, om, ov = params.copy(), np.zeros_like(params), np.zeros_like(params)
    for _ in range(3):
        ost, _ = oupd.ppo_update(osp, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, ent_coef=0.01,
                                 world=world)
    assert np.array_equal(dp.cpu().numpy(), op), f"rank {rank}: sharded update differs from the oracle"
    assert np.array_equal(st.cpu().numpy(), ost), f"rank {rank}: sharded update statistics differ from the oracle"
    dist.barrier()


def engine_checks(rank, local, world, p2p=True, env_kind="liar"):
    from pantheonrl_b200 import _lib
    from pantheonrl_b200.engine import PPOConfig, VecTrainer
    N, T = 512, 16
    cfg = PPOConfig(n_steps=T, n_epochs=2, n_minibatches=4)
    modes = [("nccl", "replicated"), ("nccl", "sharded")] + ([("p2p", "sharded")] if p2p else [])
    results = {}
    for mode in modes:
        kw = {"layout": "simple", "horizon": 12} if env_kind == "overcooked" else {}
        tr = VecTrainer(env_kind, N, cfg, seed=10, partner="ppo", device=f"cuda:{local}", env0=rank * N,
                        group=dist.group.WORLD, exchange=mode[0], ego_update=mode[1], **kw)
        for _ in range(3):
            tr.iteration()
        torch.cuda.synchronize()
        # 1. ego replicas identical on every rank; partners are different learners
        ego = [torch.empty_like(tr.ego.params) for _ in range(world)]
        dist.all_gather(ego, tr.ego.params)
        alt = [torch.empty_like(tr.alt.params) for _ in range(world)]
        dist.all_gather(alt, tr.alt.params)
        assert all(torch.equal(ego[0], e) for e in ego), f"[{mode}] ego replicas diverged"
        assert not torch.equal(alt[0], alt[1]), f"[{mode}] partners should differ"
        # 2. my slice of the gathered stream == my own records; the whole stream == all ranks' records
        count, rb = T * N, tr.rec_bytes
        lib, ctx = _lib.load(), _lib.Context.get(local)
        mine = torch.empty(count * rb, dtype=torch.uint8, device=f"cuda:{local}")
        tr.pack_into(mine)
        torch.cuda.synchronize()
        assert torch.equal(tr.gather[rank * count * rb:(rank + 1) * count * rb], mine), f"[{mode}] own slice differs"
        full = torch.empty(world * count * rb, dtype=torch.uint8, device=f"cuda:{local}")
        dist.all_gather_into_tensor(full, mine)
        assert torch.equal(tr.gather[: world * count * rb], full), f"[{mode}] gathered stream differs"
        results[mode] = (tr.ego.params.clone(), tr.alt.params.clone())
        del tr, lib, ctx
    a, b = results[("nccl", "replicated")], results[("nccl", "sharded")]
    assert torch.isfinite(a[0]).all() and torch.isfinite(b[0]).all()
    if p2p:
        a, b = results[("nccl", "sharded")], results[("p2p", "sharded")]
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), "p2p and nccl exchange lead to different parameters"
    dist.barrier()


def engine_vs_oracle(rank, local, world):
    """One iteration of the sharded engine against the oracle pieces: this rank's rollout (global env
    ids), GAE, and the ego update over the gathered global batch with the oracle's world-aware
    reduction; the partner's update over this rank's ragged buffer."""
    import oracle
    from oracle import rollout as orc, update as oupd
    from pantheonrl_b200 import _lib, update as dupd
    from pantheonrl_b200.engine import PPOConfig, VecTrainer
    N, T, E, NMB, seed = 256, 8, 2, 4, 10
    cfg = PPOConfig(n_steps=T, n_epochs=E, n_minibatches=NMB)
    tr = VecTrainer("liar", N, cfg, seed=seed, partner="ppo", device=f"cuda:{local}", env0=rank * N,
                    group=dist.group.WORLD)
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    pe, pa = tr.ego.params.cpu().numpy().copy(), tr.alt.params.cpu().numpy().copy()
    tr.iteration()
    torch.cuda.synchronize()
    # every rank's rollout on the oracle (the ego update needs all of them)
    shards = [orc.rollout("liar", osp, pe, pa, N=N, T=T, seed=seed, env0=r * N) for r in range(world)]
    o_ego, o_alt, carry = shards[rank]
    for k in ("obs", "actions", "rewards", "values", "logp"):
        assert np.array_equal(getattr(tr.ego_buf, k).cpu().numpy(), o_ego[k]), (rank, k)
    cat = {k: [] for k in ("obs", "actions", "logp", "adv", "ret")}
    for e, _, c in shards:
        adv, ret = oracle.gae(e["rewards"], e["values"], e["episode_starts"], c["ego_last_value"], c["ego_last_done"])
        for k, v in (("obs", e["obs"]), ("actions", e["actions"]), ("logp", e["logp"]), ("adv", adv), ("ret", ret)):
            cat[k].append(v.reshape(T * N, -1) if v.ndim == 3 else v.reshape(T * N))
    cat = {k: np.concatenate(v) for k, v in cat.items()}
    from pantheonrl_b200.dist_util import global_env_major_index
    idx = global_env_major_index(world, T, N).numpy()
    M = idx.size
    bs = -(-M // NMB)
    perm = oupd.perm_feistel(M, E, seed, _lib.STREAM_SHUFFLE_EGO, epoch0=0)
    me, ve = np.zeros_like(pe), np.zeros_like(pe)
    oupd.ppo_update(osp, pe, me, ve, 0, cat["obs"], cat["actions"], cat["logp"], cat["adv"], cat["ret"], perm, bs,
                    tr.last_grids[0], index=idx, world=world)
    assert np.array_equal(tr.ego.params.cpu().numpy(), pe), f"rank {rank}: sharded ego update differs from the oracle"
    aadv, aret = oracle.gae_ragged(o_alt["rewards"], o_alt["values"], o_alt["episode_starts"], o_alt["count"],
                                   carry["alt_boot_done"])
    aidx = oupd.index_build(o_alt["count"], orc.alt_capacity("liar", T), N)
    Ma = aidx.size
    bsa = -(-Ma // NMB)
    aperm = oupd.perm_feistel(Ma, E, seed, _lib.STREAM_SHUFFLE_ALT, epoch0=0)
    ma, va = np.zeros_like(pa), np.zeros_like(pa)
    oupd.ppo_update(osp, pa, ma, va, 0, o_alt["obs"], o_alt["actions"], o_alt["logp"], aadv, aret, aperm, bsa,
                    tr.last_grids[1] or dupd.update_grid(tr.space, Ma, bsa), index=aidx)
    assert np.array_equal(tr.alt.params.cpu().numpy(), pa), f"rank {rank}: partner update differs from the oracle"
    dist.barrier()


def run_all(rank, local, world, p2p=True):
    sharded_update_vs_oracle(rank, local, world)
    engine_vs_oracle(rank, local, world)
    engine_checks(rank, local, world, p2p)
    engine_checks(rank, local, world, p2p, env_kind="rps")
    engine_checks(rank, local, world, p2p, env_kind="overcooked")  # Box records (272 bytes)


def spawn_entry(rank, world, port, p2p, out):
    """torch.multiprocessing.spawn target."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        run_all(rank, rank, world, p2p)
        if rank == 0:
            open(out, "w").write(f"ok world={world}\n")
    finally:
        dist.destroy_process_group()
