"""Experiment: ego and partner updates as two CONCURRENT cooperative launches on two streams
(each pinned to half of the SMs) vs back to back on the full device."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200 import _lib, update as up
from pantheonrl_b200.engine import VecTrainer, PPOConfig

GRID = {"g": 0}
_orig = up.ppo_update
def patched(*a, **k):
    k["grid_ctas"] = GRID["g"]
    return _orig(*a, **k)
up.ppo_update = patched

def main(env="liar", N=4096, T=128):
    cfg = PPOConfig(n_steps=T, n_minibatches=32, n_epochs=10)
    tr = VecTrainer(env, N, cfg, seed=10, partner="ppo")
    for _ in range(2):
        tr.iteration()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    def prep():
        tr.collect(); tr.compute_gae()
        a = tr.alt_buf
        index, total = up.index_build(a.count, a.Tcap, tr.N, device=tr.device)
        M = int(total.item())
        torch.cuda.synchronize()
        return index, M
    def seq(g):
        index, M = prep()
        GRID["g"] = g
        e0, e1 = ev(), ev()
        e0.record()
        tr._train_one(tr.ego, tr.ego_buf, tr.ego_index, tr.ego_M, tr.ego_perm, tr.ego_ws, _lib.STREAM_SHUFFLE_EGO)
        perm = tr.alt_perm_store[: cfg.n_epochs * M].view(cfg.n_epochs, M)
        tr._train_one(tr.alt, tr.alt_buf, index, M, perm, tr.alt_ws, _lib.STREAM_SHUFFLE_ALT)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def conc(g):
        index, M = prep()
        GRID["g"] = g
        e0, e1 = ev(), ev()
        cur = torch.cuda.current_stream()
        e0.record()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            tr._train_one(tr.ego, tr.ego_buf, tr.ego_index, tr.ego_M, tr.ego_perm, tr.ego_ws, _lib.STREAM_SHUFFLE_EGO)
        with torch.cuda.stream(s2):
            perm = tr.alt_perm_store[: cfg.n_epochs * M].view(cfg.n_epochs, M)
            tr._train_one(tr.alt, tr.alt_buf, index, M, perm, tr.alt_ws, _lib.STREAM_SHUFFLE_ALT)
        cur.wait_stream(s1); cur.wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    for g in (0, 74, 64):
        print(f"sequential grid={g or 'auto'}: {seq(g):.2f} ms  {seq(g):.2f} ms")
    for g in (74, 64, 48):
        print(f"concurrent grid={g} each: {conc(g):.2f} ms  {conc(g):.2f} ms")

if __name__ == "__main__":
    main(*(sys.argv[1:2]), *map(int, sys.argv[2:4]))
