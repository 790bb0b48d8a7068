"""One short trainer iteration for ncu captures (dev tool)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200.engine import VecTrainer, PPOConfig

env = sys.argv[1] if len(sys.argv) > 1 else "liar"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 128
epochs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
cfg = PPOConfig(n_steps=T, n_minibatches=32, n_epochs=epochs)
tr = VecTrainer(env, N, cfg, seed=10, partner="ppo")
for _ in range(2):
    tr.iteration()
torch.cuda.synchronize()
print("ok", tr.train_stats())
