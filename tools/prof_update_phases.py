"""Phase timeline of the update kernel (dev tool): CTA 0's clock64 sums per phase,
averaged per minibatch.  usage: python tools/prof_update_phases.py [env N T]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200 import _lib, update as up
from pantheonrl_b200.engine import VecTrainer, PPOConfig

GRID = int(os.environ.get("PTH_GRID", "0"))  # pin the update grid (0 = auto)
_orig = up.ppo_update
up.ppo_update = lambda *a, **k: _orig(*a, **{**k, "grid_ctas": GRID})

NAMES = ["loop head", "weights->smem", "gather", "pi L0", "pi L1", "head+loss", "head wgrad|dz2",
         "pi tower bwd", "vf L0", "vf L1", "value head", "vf tower bwd", "tile stats", "barrier1",
         "reduce", "barrier2", "adam", "barrier3", " pi wgrad64", " pi backprop64", " vf wgrad64", " vf backprop64", "slot sort+rows", "setup"]


def run(env, N, T, **kw):
    cfg = PPOConfig(n_steps=T, n_minibatches=32, n_epochs=10)
    tr = VecTrainer(env, N, cfg, seed=10, partner="ppo", **kw)
    for _ in range(2):
        tr.iteration()
    tr.collect(); tr.compute_gae()
    prof = torch.zeros(32, dtype=torch.int64, device="cuda")
    _lib.load().pth_debug_update_profile(prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.train(); e1.record()
    torch.cuda.synchronize()
    _lib.load().pth_debug_update_profile(None)
    ms = e0.elapsed_time(e1)
    p = prof.cpu().numpy().astype(float)
    n_mb = 2 * 10 * 32
    tot = p.sum()
    print(f"== {env} N={N} T={T}: train {ms:.2f} ms, {ms / n_mb * 1e3:.1f} us / minibatch; clock sum {tot / n_mb:.0f} cyc / minibatch")
    for i, nm in enumerate(NAMES):
        print(f"  {nm:16s} {p[i] / n_mb:9.0f} cyc  {100 * p[i] / tot:5.1f}%  ~{p[i] / tot * ms / n_mb * 1e3:6.2f} us")


if __name__ == "__main__":
    if len(sys.argv) > 3:
        run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
    else:
        run("liar", 4096, 128)
        run("rps", 4096, 128)
        run("overcooked", 1024, 400, layout="simple")
