"""torchrun --nproc-per-node W tools/prof_update_phases_mgpu.py
Phase timeline (CTA 0 of rank 0) of the SHARDED ego update alone (dev tool)."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200 import _lib
from pantheonrl_b200.engine import VecTrainer, PPOConfig
from prof_update_phases import NAMES

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, T = 4096, 128
cfg = PPOConfig(n_steps=T, n_minibatches=32, n_epochs=10)
tr = VecTrainer("liar", N, cfg, seed=10, partner="ppo", device=f"cuda:{local}", env0=rank * N,
                group=dist.group.WORLD, concurrent_updates=False)
for _ in range(2):
    tr.iteration()
tr.collect(); tr.compute_gae(); tr.exchange_ego()
prof = torch.zeros(32, dtype=torch.int64, device="cuda")
if rank == 0:
    _lib.load().pth_debug_update_profile(prof.data_ptr())
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr._train_one(tr.ego, tr.ego_buf, tr.ego_index, tr.ego_M, tr.ego_perm, tr.ego_ws, _lib.STREAM_SHUFFLE_EGO,
              packed=tr.gather, peers=tr.peers)
e1.record(); torch.cuda.synchronize()
_lib.load().pth_debug_update_profile(None)
if rank == 0:
    ms = e0.elapsed_time(e1)
    p = prof.cpu().numpy().astype(float)
    n_mb, tot = 320, p.sum()
    print(f"== sharded ego update, world={world}: {ms:.2f} ms, {ms / n_mb * 1e3:.1f} us / minibatch")
    for i, nm in enumerate(NAMES):
        print(f"  {nm:16s} {p[i] / n_mb:9.0f} cyc  {100 * p[i] / tot:5.1f}%  ~{p[i] / tot * ms / n_mb * 1e3:6.2f} us")
dist.barrier()
dist.destroy_process_group()
