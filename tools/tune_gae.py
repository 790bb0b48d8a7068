"""Sweep GAE kernel variants / tuning codes on the GPU box; writes gpurun_out/tune_gae.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pantheonrl_b200 import ops  # noqa: E402


def time_variant(args, variant, iters=10, flush=None):
    adv = torch.empty_like(args[0])
    ret = torch.empty_like(args[0])
    for _ in range(3):
        ops.gae(*args, variant=variant, out=(adv, ret))
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gae(*args, variant=variant, out=(adv, ret))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    out = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (T, N) in [(2048, 65536), (128, 65536), (128, 4096), (2048, 4096)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        rew = torch.randint(-1, 2, (T, N), generator=g, device="cuda").float()
        val = torch.randn(T, N, generator=g, device="cuda")
        start = (torch.rand(T, N, generator=g, device="cuda") < 0.25).float()
        lv = torch.randn(N, generator=g, device="cuda")
        dn = (torch.rand(N, generator=g, device="cuda") < 0.25).float()
        args = (rew, val, start, lv, dn)
        nbytes = 20 * T * N + 8 * N
        variants = [("auto", 0), ("ldg", 1), ("tma", 2)]
        for vec in (1, 2, 4):
            for uc, U in ((1, 4), (2, 8), (3, 16)):
                for blk in (1, 2, 4):
                    variants.append((f"ldg v{vec} u{U} b{blk*32}", 0x1000 | (vec << 8) | (uc << 4) | blk))
        for per_sm in (1, 2, 3, 4):
            for ttc, tt in ((4, 2), (1, 4), (2, 8), (3, 16)):
                for st in (2, 3, 4, 6, 8, 12):
                    if (tt, st) not in ((2, 8), (2, 12), (4, 2), (4, 4), (4, 6), (4, 8), (4, 12), (8, 2), (8, 3),
                                        (8, 4), (8, 6), (16, 2), (16, 3), (16, 4)):
                        continue
                    variants.append((f"tma sm{per_sm} tt{tt} s{st}", 0x2000 | (per_sm << 8) | (ttc << 4) | st))
        for name, v in variants:
            try:
                med, best = time_variant(args, v, flush=flush)
            except Exception as e:  # noqa: BLE001
                out.append(dict(T=T, N=N, name=name, error=str(e)[:120]))
                continue
            out.append(dict(T=T, N=N, name=name, variant=v, ms=med, ms_best=best,
                            gbs=nbytes / med / 1e6, gbs_best=nbytes / best / 1e6))
        del rew, val, start
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/tune_gae.json", "w"), indent=1)
    for (T, N) in sorted({(o["T"], o["N"]) for o in out}):
        rows = sorted([o for o in out if (o["T"], o["N"]) == (T, N) and "gbs" in o], key=lambda o: -o["gbs"])
        print(f"T={T} N={N}")
        for o in rows[:6]:
            print(f"   {o['name']:24s} {o['ms']:.4f} ms  {o['gbs']:.0f} GB/s (best {o['gbs_best']:.0f})")
        for o in rows:
            if o["name"] in ("auto", "ldg", "tma"):
                print(f"   [{o['name']}] {o['ms']:.4f} ms  {o['gbs']:.0f} GB/s")


if __name__ == "__main__":
    main()
