"""torchrun --nproc-per-node G tools/multi_gpu_check.py [p2p]
Multi-GPU parity of the one-partner-per-GPU engine: ego replicas stay bit-identical,
partners differ, the gathered stream equals every rank's own packed records, and
(optionally) the fused pack+P2P all-gather kernel equals pack + ncclAllGather."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200 import _lib
from pantheonrl_b200.engine import PPOConfig, VecTrainer


def sharded_update_vs_oracle(rank, local, world):
    """pth_ppo_update with world > 1 (tiles sharded over ranks, in-kernel peer exchange of
    the per-rank gradient sums) against the CPU oracle's world-aware reduction, bit for bit."""
    import numpy as np
    import oracle
    from oracle import update as oupd
    from pantheonrl_b200 import update as dupd, rollout as ro
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_oracle_update import make_batch
    from pantheonrl_b200.policy import init_flat
    kw = oracle.LIAR_SPACE
    sp, osp = ro.space_for("liar"), oracle.make_space(**kw)
    M, BS, E, grid = 1500, 640, 2, 2
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
    for _ in range(2):  # two launches: the flag epoch carries over
        st = dupd.ppo_update(sp, dp, dm, dv, 0, d(obs), d(act), d(old_logp), d(adv), d(ret), d(perm), BS, ws,
                             grid_ctas=grid, peers=peers, ent_coef=0.01)
    torch.cuda.synchronize()
    op, om, ov = params.copy(), np.zeros_like(params), np.zeros_like(params)
    for _ in range(2):
        ost, _ = oupd.ppo_update(osp, op, om, ov, 0, obs, act, old_logp, adv, ret, perm, BS, grid, ent_coef=0.01,
                                 world=world)
    assert np.array_equal(dp.cpu().numpy(), op), f"rank {rank}: sharded update differs from the oracle"
    assert np.array_equal(st.cpu().numpy(), ost)
    dist.barrier()
    if rank == 0:
        print(f"sharded pth_ppo_update (world={world}) bit-exact vs oracle")


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, T = 512, 16
    cfg = PPOConfig(n_steps=T, n_epochs=2, n_minibatches=4)
    modes = [("nccl", "replicated"), ("nccl", "sharded")] + ([("p2p", "sharded")] if "p2p" in sys.argv else [])
    sharded_update_vs_oracle(rank, local, world)
    results = {}
    for mode in modes:
        tr = VecTrainer("liar", N, cfg, seed=10, partner="ppo", device=f"cuda:{local}", env0=rank * N,
                        group=dist.group.WORLD, exchange=mode[0], ego_update=mode[1])
        for _ in range(2):
            tr.iteration()
        torch.cuda.synchronize()
        # 1. ego replicas identical on every rank; partners are different learners
        ego = [torch.empty_like(tr.ego.params) for _ in range(world)]
        dist.all_gather(ego, tr.ego.params)
        alt = [torch.empty_like(tr.alt.params) for _ in range(world)]
        dist.all_gather(alt, tr.alt.params)
        assert all(torch.equal(ego[0], e) for e in ego), f"[{mode}] ego replicas diverged"
        assert world == 1 or not torch.equal(alt[0], alt[1]), f"[{mode}] partners should differ"
        # 2. my slice of the gathered stream == my own records
        count = T * N
        lib, ctx = _lib.load(), _lib.Context.get(local)
        mine = torch.empty(count * 48, dtype=torch.uint8, device=f"cuda:{local}")
        b = tr.ego_buf
        _lib.check(lib.pth_pack_transitions(ctx.handle, b.obs.data_ptr(), b.actions.data_ptr(), b.logp.data_ptr(),
                                            b.advantages.data_ptr(), b.returns.data_ptr(), count, mine.data_ptr(),
                                            _lib.current_stream()), "pack")
        torch.cuda.synchronize()
        assert torch.equal(tr.gather[rank * count * 48:(rank + 1) * count * 48], mine), f"[{mode}] own slice differs"
        results_key = mode
        full = torch.empty(world * count * 48, dtype=torch.uint8, device=f"cuda:{local}")
        dist.all_gather_into_tensor(full, mine)
        assert torch.equal(tr.gather[: world * count * 48], full), f"[{mode}] gathered stream differs"
        results[mode] = (tr.ego.params.clone(), tr.alt.params.clone())
        del tr
    if ("p2p", "sharded") in results:
        a, b = results[("nccl", "sharded")], results[("p2p", "sharded")]
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), "p2p and nccl exchange lead to different parameters"
    dist.barrier()
    if rank == 0:
        print(f"multi_gpu_check ok: world={world} modes={modes}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
