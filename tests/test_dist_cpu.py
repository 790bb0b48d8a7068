"""CPU tests (gloo, world_size 2) of the N > 1 host logic: env sharding by global
env id, the 48-byte exchange record, the all-gather and the global env-major
sample index.  The per-rank compute is the CPU oracle; the claim checked is that
two sharded ranks + one all-gather reproduce a single process over all envs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from oracle import rollout as orc
from pantheonrl_b200 import dist_util

N, T, SEED = 24, 10, 5


def pack_records(ego, adv, ret):
    """numpy restatement of pth_pack_transitions' record layout."""
    cnt = ego["rewards"].size
    rec = np.zeros((cnt, dist_util.PACKED_BYTES), np.uint8)
    rec[:, 0:32] = ego["obs"].reshape(cnt, 32)
    rec[:, 32:36] = ego["actions"].reshape(cnt, 4)
    rec[:, 36:40] = ego["logp"].reshape(cnt).view(np.uint8).reshape(cnt, 4)
    rec[:, 40:44] = adv.reshape(cnt).view(np.uint8).reshape(cnt, 4)
    rec[:, 44:48] = ret.reshape(cnt).view(np.uint8).reshape(cnt, 4)
    return rec


def one_rank(env0, n):
    osp = oracle.make_space(**oracle.LIAR_SPACE)
    P = oracle.param_count(osp)
    pe = (np.random.RandomState(1).randn(P) * 0.2).astype(np.float32)
    pa = (np.random.RandomState(2 + env0).randn(P) * 0.2).astype(np.float32)  # each rank has its own partner
    ego, alt, carry = orc.rollout("liar", osp, pe, pa, N=n, T=T, seed=SEED, env0=env0)
    adv, ret = oracle.gae(ego["rewards"], ego["values"], ego["episode_starts"], carry["ego_last_value"],
                          carry["ego_last_done"])
    return ego, adv, ret


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ego, adv, ret = one_rank(dist_util.shard_env0(rank, N), N)
    mine = torch.from_numpy(pack_records(ego, adv, ret).reshape(-1))
    gathered = torch.empty(world * mine.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, mine)
    if rank == 0:
        np.save(out, gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_sharded_ranks_equal_one_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    gathered = np.load(out).reshape(-1, dist_util.PACKED_BYTES)
    index = dist_util.global_env_major_index(2, T, N).numpy()
    assert np.array_equal(np.sort(index), np.arange(2 * T * N))
    stream = gathered[index]  # what the update kernel reads, in sample order

    # single process over the union of envs, with the matching partner per shard
    parts = [one_rank(0, N), one_rank(N, N)]
    for field, lo, hi, src in (("obs", 0, 32, "obs"), ("actions", 32, 36, "actions")):
        want = np.concatenate([np.swapaxes(p[0][src], 0, 1).reshape(-1, hi - lo) for p in parts])
        assert np.array_equal(stream[:, lo:hi], want), field
    want_adv = np.concatenate([p[1].T.reshape(-1) for p in parts])
    assert np.array_equal(stream[:, 40:44].copy().view(np.float32).reshape(-1), want_adv)
    # the ego stream of global env g does not depend on how envs are sharded:
    # (same ego weights, same global env id, same partner) -> same trajectory
    whole_first = one_rank(0, N // 2)
    assert np.array_equal(whole_first[0]["actions"], parts[0][0]["actions"][:, : N // 2])


def test_global_index_layout():
    idx = dist_util.global_env_major_index(2, 3, 2).tolist()
    # rank 0: envs 0,1 ; rank 1: envs 2,3 ; record offset = r*6 + t*2 + n ; order env-major
    assert idx == [0, 2, 4, 1, 3, 5, 6, 8, 10, 7, 9, 11]
    assert dist_util.shard_env0(3, 4096) == 12288


def test_update_grid_split_is_rank_independent_for_the_ego():
    """Replicated / sharded ego updates need the same grid on every rank although the ranks'
    partner batches differ (ragged): the ego's share may not depend on the partner's tiles."""
    from pantheonrl_b200.engine import split_update_grids
    for cap in (148, 132, 16):
        for te in (1, 6, 64, 128, 129, 1000):
            shares = {split_update_grids(cap, te, ta) for ta in (0, 1, 7, 128, 129, 130, 257, 5000)}
            assert len({ge for ge, _ in shares}) == 1, (cap, te, shares)
            for ge, ga in shares:
                assert 1 <= ge and ge + ga <= cap and (ga == 0 or ga >= 1)
    assert split_update_grids(148, 128, 129) == (64, 65)   # BASELINE configs[1]: two rounds each
    assert split_update_grids(148, 1, 1) == (48, 48)       # n_envs = 1: spare CTAs share reduction + Adam
