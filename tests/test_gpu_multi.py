"""Multi-GPU parity as tests (one process per GPU, NCCL): skipped on a single-GPU box.  The checks
themselves are in tests/mgpu_worker.py; tools/multi_gpu_check.py runs the same under torchrun."""
import os
import socket
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_engine_parity(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import mgpu_worker
    out = str(tmp_path / "ok.txt")
    mp.spawn(mgpu_worker.spawn_entry, args=(world, _free_port(), True, out), nprocs=world, join=True)
    assert open(out).read().startswith("ok")
