"""Launches of the ADAP and ModularAlgorithm variants of the update kernel for ncu captures (dev tool):
ADAP.train and ModularAlgorithm.train (2 partners) at the reference's shape (LiarsDice-v0, n_steps 2048,
batch_size 64, 10 epochs), through bench.variant_config (product code only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

for kind in ("adap", "modular"):
    print(kind, bench.variant_config(torch, kind, cpu=False))
