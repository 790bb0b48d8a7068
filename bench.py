#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native PPO rollout/GAE/update path.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle port of the
                                                            reference's path on the host cores)

A "step" is one full train iteration of BASELINE.json configs[1]
(LiarsDice-v0 PPO-vs-PPO, 4096 on-device envs per GPU, n_steps=128, 10 epochs x
32 minibatches, SB3 default hyper-parameters): rollout of both agents + GAE +
PPO.train of both learners.  metric = agent decisions (ego + partner) per second.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (all agents)"
UNIT = "agent-steps/s"
N_ENVS, N_STEPS, N_EPOCHS, N_MB = 4096, 128, 10, 32
FWD_FLOPS = {"liar": 88064, "rps": 17152}  # dense-convention forward FLOPs per sample


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--env", default="liar", choices=["liar", "rps"])
    ap.add_argument("--n-envs", type=int, default=N_ENVS)
    ap.add_argument("--cpu-sample-envs", type=int, default=0, help="0 = same env count as the GPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"])
    ap.add_argument("--ego-update", default="sharded", choices=["sharded", "replicated"])
    return ap.parse_args()


# ----------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------- CPU arm
CPU_BUDGET_S = 150.0   # the reference arm stops adding steps after this much wall time
CPU_EPOCHS_TIMED = 1  # of N_EPOCHS: the update time of the sample is scaled by N_EPOCHS / this


def cpu_iteration_rate(env, n_envs, steps, warmup):
    """agent-steps/s of the oracle port (oracle/cpu_trainer.py) on the host cores.

    Bounded sample of the GPU arm's step: the same iteration (T = 128, 32 minibatches
    per epoch, both learners) at n_envs envs, with CPU_EPOCHS_TIMED of the 10 epochs
    actually run and the update time scaled up to 10 (epochs are identical work)."""
    import torch
    from oracle.cpu_trainer import CpuTrainer
    cores = os.cpu_count() or 1
    # torch CPU eager over-subscribes badly on tiny ops: cap its intra-op pool
    threads = min(cores, 16)
    torch.set_num_threads(threads)
    tr = CpuTrainer(env, n_envs, n_steps=N_STEPS, n_epochs=CPU_EPOCHS_TIMED, n_minibatches=N_MB, seed=10)
    for _ in range(warmup):
        tr.iteration()
    dec, total, phases = 0, 0.0, {"rollout_s": 0.0, "gae_s": 0.0, "train_s": 0.0}
    t_begin, done = time.perf_counter(), 0
    for _ in range(steps):
        dec += tr.iteration()
        done += 1
        t = tr.timing
        total += t["rollout_s"] + t["gae_s"] + t["train_s"] * (N_EPOCHS / CPU_EPOCHS_TIMED)
        for k in phases:
            phases[k] += t[k]
        if time.perf_counter() - t_begin > CPU_BUDGET_S:  # bounded: stop early, report the steps really run
            break
    phases = {k: v / done for k, v in phases.items()}
    phases["train_s_scaled_to_10_epochs"] = phases["train_s"] * N_EPOCHS / CPU_EPOCHS_TIMED
    return dec / total, total / done, threads, phases, done


def cpu_sample_text(args):
    return (f"per step: one train iteration (rollout + GAE + PPO.train of both learners, T={N_STEPS}, "
            f"{N_MB} minibatches/epoch) at {args.cpu_sample_envs} envs (GPU arm: {args.n_envs}); "
            f"{CPU_EPOCHS_TIMED} of {N_EPOCHS} epochs run, update time scaled x{N_EPOCHS // CPU_EPOCHS_TIMED} "
            f"(epochs are identical work); OpenMP C rollout + torch CPU eager update (threads capped at 16: "
            f"more is slower)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args.cpu_sample_envs = args.cpu_sample_envs or args.n_envs
    rate, spi, cores, timing, done = cpu_iteration_rate(args.env, args.cpu_sample_envs, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": spi * 1e3,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.env}-ppo-vs-ppo", "n_envs_per_gpu": args.n_envs, "n_steps": N_STEPS,
                   "n_epochs": N_EPOCHS, "n_minibatches": N_MB},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(args), "phase_seconds": timing,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args):
    """The CPU leg runs in a child with a hard timeout so it can never stall the GPU line."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "0",
           "--env", args.env, "--n-envs", str(args.n_envs), "--cpu-sample-envs", str(args.cpu_sample_envs)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240).stdout.strip().splitlines()
        return json.loads(out[-1])["cpu_baseline"]
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                "sample": f"CPU leg failed or exceeded 240 s: {type(e).__name__}"}


# ----------------------------------------------------------------------- GPU arm
def gae_roofline(torch, ops, peaks):
    """GAE kernel at BASELINE configs[2] size; inputs (2.7 GB) exceed the 126 MB L2."""
    T, N = 2048, 65536
    g = torch.Generator(device="cuda").manual_seed(0)
    rew = torch.randint(-1, 2, (T, N), generator=g, device="cuda").float()
    val = torch.randn(T, N, generator=g, device="cuda")
    start = (torch.rand(T, N, generator=g, device="cuda") < 0.25).float()
    lv = torch.randn(N, generator=g, device="cuda")
    dn = (torch.rand(N, generator=g, device="cuda") < 0.25).float()
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    for _ in range(3):
        ops.gae(rew, val, start, lv, dn, out=(adv, ret))
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gae(rew, val, start, lv, dn, out=(adv, ret))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 20 * T * N + 8 * N
    achieved = nbytes / ms / 1e6
    peak = peaks.get("hbm_gbs")
    src = "measured (MEASURED_PEAKS.json hbm_gbs, burst)"
    if not peak:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    del rew, val, start, adv, ret
    torch.cuda.empty_cache()
    return {"kernel": "gae_tma_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": src, "shape": [T, N], "bytes_per_launch": nbytes,
            "ms_per_launch": ms, "traffic": 2.69e9,
            "traffic_source": "profiles/gae_r01.md (ncu dram__bytes_read+write per launch)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pantheonrl_b200 import _lib, ops
    from pantheonrl_b200.engine import PPOConfig, VecTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # the contract is ONE stdout line; NCCL_DEBUG=VERSION would add one
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass

    cfg = PPOConfig(n_steps=N_STEPS, n_epochs=N_EPOCHS, n_minibatches=N_MB)
    # one partner per GPU, envs sharded: rank r owns global envs [r*N, (r+1)*N)
    tr = VecTrainer(args.env, args.n_envs, cfg, seed=10, partner="ppo", device=f"cuda:{local}",
                    env0=rank * args.n_envs, group=dist.group.WORLD if world > 1 else None,
                    exchange=args.exchange, ego_update=args.ego_update)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        tr.iteration()
    barrier()

    # ---- timed region: K iterations, device timing, phases on the launching stream
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks = []
    launches0 = _lib.LAUNCHES
    decisions = 0
    # L2 flush between timed steps: 256 MB written (> the 126 MB L2), outside the per-step event brackets
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with ClockSampler(local) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)
            a, b, c, d = ev(), ev(), ev(), ev()
            a.record()
            tr.collect()
            b.record()
            tr.compute_gae()
            c.record()
            m_alt = tr.train()
            d.record()
            marks.append((a, b, c, d))
            decisions += tr.N * tr.T + m_alt
        barrier()
        t_wall = time.perf_counter() - t_wall0
    clocks = clk.summary()
    launches = _lib.LAUNCHES - launches0
    ms_total = sum(a.elapsed_time(d) for a, b, c, d in marks)  # K steps, device time, flushes excluded
    ph = {"rollout_ms": sum(a.elapsed_time(b) for a, b, c, d in marks) / args.steps,
          "gae_ms": sum(b.elapsed_time(c) for a, b, c, d in marks) / args.steps,
          "update_ms": sum(c.elapsed_time(d) for a, b, c, d in marks) / args.steps}

    # max over ranks of the device time; sum of decisions
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        n = torch.tensor([float(decisions)], device="cuda", dtype=torch.float64)
        dist.all_reduce(n)
        decisions = float(n.item())
    value = decisions / (ms_total / 1e3)

    # ---- e2e: same iterations through the public API with HOST-resident learner state:
    # every step uploads both learners' parameters + Adam moments from pinned host
    # memory, runs the iteration, downloads the updated state and the logged scalars.
    host = {}
    for name, L in (("ego", tr.ego), ("alt", tr.alt)):
        host[name] = [t.cpu().pin_memory() for t in (L.params, L.adam_m, L.adam_v)]
    stats_host = torch.empty_like(tr.ego.last_stats, device="cpu").pin_memory()
    h2d = sum(t.numel() * 4 for v in host.values() for t in v)
    d2h = h2d + stats_host.numel() * 4 * 2 + 16
    barrier()
    t0 = time.perf_counter()
    e2e_dec = 0
    for _ in range(args.steps):
        for name, L in (("ego", tr.ego), ("alt", tr.alt)):
            for dst, src in zip((L.params, L.adam_m, L.adam_v), host[name]):
                dst.copy_(src, non_blocking=True)
        e2e_dec += tr.iteration()
        for name, L in (("ego", tr.ego), ("alt", tr.alt)):
            for src, dst in zip((L.params, L.adam_m, L.adam_v), host[name]):
                dst.copy_(src, non_blocking=True)
        stats_host.copy_(tr.ego.last_stats, non_blocking=True)
        _ = tr.alt.last_stats.cpu()
        _ = tr.carry.ep_stats.cpu()
        torch.cuda.synchronize()
    barrier()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
        n = torch.tensor([float(e2e_dec)], device="cuda", dtype=torch.float64)
        dist.all_reduce(n)
        e2e_dec = float(n.item())
    e2e = {"value": e2e_dec / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "api": "VecTrainer.iteration() with host-resident (pinned) learner state uploaded/downloaded every step"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines (rank 0, N = 1 only for the standalone GAE measurement)
    roof = gae_roofline(torch, ops, peaks) if world == 1 else None
    M_ego = tr.N * tr.T
    flops = 3 * FWD_FLOPS[args.env] * N_EPOCHS * (M_ego + decisions / args.steps / world - M_ego)
    sm_mhz = clocks.get("sm_mhz") or 1500.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    dom = {"kernel": "ppo_update_kernel (x2 learners)", "bound": "fp32-ffma (not hbm/tensor: parity mode is fp32 CUDA-core math)",
           "achieved": flops / (ph["update_ms"] / 1e3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
           "frac": flops / (ph["update_ms"] / 1e3) / 1e12 / fp32_peak,
           "peak_source": "148 SMs x 128 FFMA/clk x 2 x median SM clock under load",
           "flops_convention": "dense: 3 x forward FLOPs per sample per epoch (SURVEY.md 8d)"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(args)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.env}-ppo-vs-ppo (BASELINE configs[1] per GPU; one partner per GPU)",
                   "n_envs_per_gpu": args.n_envs, "n_steps": N_STEPS, "n_epochs": N_EPOCHS,
                   "n_minibatches": N_MB, "partners": world, "parallelism": f"dp{world}: envs + one partner per GPU, "
                   f"ego replicated, 1 all-gather of packed ego transitions per rollout ({args.exchange}), ego update "
                   f"{args.ego_update if world > 1 else 'local'}", "l2": "flushed between timed steps (256 MB "
                   "fill outside the per-step CUDA-event brackets); within a step the rollout buffers (~90 MB) are "
                   "produced and consumed on the device"},
        "phases_ms": ph, "wall_s": t_wall, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roof, "roofline_dominant": dom, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
