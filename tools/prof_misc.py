"""Launches of the smaller kernels for ncu captures (dev tool): policy_forward_kernel (B = 65 536, Liar's
Dice space), pack_kernel (524 288 one-hot records and 409 600 Box records), gae_ragged_kernel (partner buffer
of the bench configuration)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200 import _lib, ops, rollout as ro  # noqa: E402
from pantheonrl_b200.engine import PPOConfig, VecTrainer  # noqa: E402
from pantheonrl_b200.policy import init_flat  # noqa: E402

sp = ro.space_for("liar")
params = torch.from_numpy(init_flat(sp, 10)).cuda()
B = 65536
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.zeros(B, 32, dtype=torch.uint8, device="cuda")
obs[:, :6] = torch.randint(0, 7, (B, 6), generator=g, device="cuda", dtype=torch.uint8)
obs[:, 6:30:2] = 6
for _ in range(3):
    out = ops.policy_forward(sp, params, obs, seed=1, want=("action", "value", "logp"))
torch.cuda.synchronize()

tr = VecTrainer("liar", 4096, PPOConfig(n_steps=128, n_epochs=1, n_minibatches=32), seed=10, partner="ppo")
tr.collect()
for _ in range(3):
    tr.compute_gae()  # gae kernel + gae_ragged_kernel
packed = torch.empty(tr.T * tr.N * tr.rec_bytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    tr.pack_into(packed)
torch.cuda.synchronize()
oc = VecTrainer("overcooked", 1024, PPOConfig(n_steps=400, n_epochs=1, n_minibatches=32), seed=10, partner="ppo",
                layout="simple")
oc.collect()
oc.compute_gae()
packed = torch.empty(oc.T * oc.N * oc.rec_bytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    oc.pack_into(packed)
torch.cuda.synchronize()
print("ok", _lib.LAUNCHES)
