"""GAE kernel at the roofline size for ncu captures (dev tool)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200 import ops
T, N = 2048, 65536
g = torch.Generator(device="cuda").manual_seed(0)
rew = torch.randint(-1, 2, (T, N), generator=g, device="cuda").float()
val = torch.randn(T, N, generator=g, device="cuda")
start = (torch.rand(T, N, generator=g, device="cuda") < 0.25).float()
lv = torch.randn(N, generator=g, device="cuda"); dn = (torch.rand(N, generator=g, device="cuda") < 0.25).float()
for v in (2, 1, 2):
    ops.gae(rew, val, start, lv, dn, variant=v)
torch.cuda.synchronize()
