"""torchrun --nproc-per-node G tools/multi_gpu_check.py [nop2p]
The multi-GPU parity checks of tests/mgpu_worker.py (also run by tests/test_gpu_multi.py) under torchrun."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mgpu_worker  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mgpu_worker.run_all(rank, local, world, p2p="nop2p" not in sys.argv)
    if rank == 0:
        print(f"multi_gpu_check ok: world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
