#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native PPO rollout/GAE/update path.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle port of the
                                                            reference's path on the host cores)

A "step" is one full train iteration of BASELINE.json configs[1]
(LiarsDice-v0 PPO-vs-PPO, 4096 on-device envs per GPU, n_steps=128, 10 epochs x
32 minibatches, SB3 default hyper-parameters): rollout of both agents + GAE +
PPO.train of both learners.  metric = agent decisions (ego + partner) per second.
Prints ONE JSON line on stdout (rank 0).

Both arms print the same `config` dict and use the same warm-up rule.  The CPU arm's step is a
bounded sample of the workload (`cpu_baseline.sample` says which): fewer envs, the SAME
minibatch size, all epochs really run — nothing is extrapolated.

At N = 1 the line also carries `configs`: the other BASELINE.json configurations measured in the
same run (facade N = 1 through PPO.learn, RPS self-play at 65 536 envs, Overcooked at 1024 envs),
each with the CPU port's number beside it.
"""
import argparse
import csv
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
N_EPOCHS, N_MB = 10, 32
FWD_FLOPS = {"liar": 88064, "rps": 17152, "overcooked": 33152}  # dense forward FLOPs per sample

# BASELINE.json configs.  cpu_*: the bounded sample the CPU port runs per step — fewer envs and
# proportionally fewer minibatches, so that the minibatch size (what torch-eager's efficiency
# depends on) is the full workload's.
WORKLOADS = {
    "liar": dict(baseline="configs[1]", env="liar", n_envs=4096, n_steps=128, partner="ppo",
                 name="LiarsDice-v0 PPO-vs-PPO, 4096 vectorised envs", cpu_envs=1024, cpu_mb=8),
    "rps_selfplay": dict(baseline="configs[2]", env="rps", n_envs=65536, n_steps=128, partner="selfplay",
                         name="RPS-v0 self-play, 65536 vectorised envs, 64-64 MLP", cpu_envs=8192, cpu_mb=4),
    "overcooked": dict(baseline="configs[3]", env="overcooked", n_envs=1024, n_steps=400, partner="ppo",
                       name="OvercookedMultiEnv-v0 layout=simple PPO-vs-PPO, 1024 envs (env on device)",
                       cpu_envs=128, cpu_mb=4),
    "rps_single": dict(baseline="configs[0]", env="rps", n_envs=1, n_steps=2048, partner="ppo",
                       name="RPS-v0 PPO-vs-PPO seed=10, 2 agents, 1 env (trainer.py shape: n_steps 2048, batch 64)",
                       cpu_envs=1, cpu_mb=0),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="liar", choices=sorted(WORKLOADS))
    ap.add_argument("--n-envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--partners", type=int, default=0,
                    help="partner learners in total (default: one per GPU); more than --gpus puts several on a GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"])
    ap.add_argument("--ego-update", default="sharded", choices=["sharded", "replicated"])
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)  # timing rule: at least 3 warm-up steps (both arms)
    return a


def make_config(args, world):
    """The `config` dict: identical in both arms."""
    w = WORKLOADS[args.workload]
    n_envs = args.n_envs or w["n_envs"]
    partners = args.partners or world
    cfg = {"workload": f"{w['name']} (BASELINE {w['baseline']}; per GPU)", "n_envs_per_gpu": n_envs,
           "n_steps": w["n_steps"], "n_epochs": N_EPOCHS, "n_minibatches": N_MB, "partners": partners if w["partner"] == "ppo" else 0,
           "parallelism": (f"dp{world}: envs sharded, {partners} partner learner(s) over {world} GPU(s), ego replicated, "
                           f"1 all-gather of packed ego transitions per rollout ({args.exchange}), ego update "
                           f"{args.ego_update if world > 1 else 'local'}"),
           "l2": "flushed between timed steps (256 MB fill outside the per-step CUDA-event brackets); within a step the "
                 "rollout buffers (~90 MB) are produced and consumed on the device"}
    return cfg


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
def cpu_rate(workload, steps, warmup):
    """agent-steps/s of the oracle port (oracle/cpu_trainer.py: OpenMP C rollout + torch CPU eager
    update, the op sequence SB3 executes) on the host cores.  Every step runs ALL epochs."""
    import torch
    from oracle.cpu_trainer import CpuTrainer
    w = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    threads = min(cores, 16)  # torch CPU eager over-subscribes on these op sizes: more threads is slower
    torch.set_num_threads(threads)
    tr = CpuTrainer(w["env"], w["cpu_envs"], n_steps=w["n_steps"], n_epochs=N_EPOCHS, n_minibatches=w["cpu_mb"],
                    batch_size=64, seed=10, partner=w["partner"])
    for _ in range(warmup):
        tr.iteration()
    dec, total, phases = 0, 0.0, {"rollout_s": 0.0, "gae_s": 0.0, "train_s": 0.0}
    for _ in range(steps):
        t0 = time.perf_counter()
        dec += tr.iteration()
        total += time.perf_counter() - t0
        for k in phases:
            phases[k] += tr.timing[k] / steps
    mb = f"{w['cpu_mb']} minibatches" if w["cpu_mb"] else "batch_size 64"
    sample = (f"per step: one full train iteration (rollout + GAE + PPO.train of every learner, T={w['n_steps']}, "
              f"all {N_EPOCHS} epochs run) at {w['cpu_envs']} envs x {mb} per epoch "
              f"(GPU arm: {w['n_envs']} envs x {N_MB} minibatches; same minibatch size); "
              f"OpenMP C rollout + torch CPU eager update, {threads} torch threads")
    return dict(value=dec / total, unit=UNIT, cores=threads, kind="port", sample=sample, phase_seconds=phases,
                host_cpus=cores, steps=steps, warmup=warmup, s_per_step=total / steps)


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cpu = cpu_rate(args.workload, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["s_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args, world), "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(workload, steps=3, warmup=1, timeout=300):
    """The CPU leg runs in a child with a hard timeout so it can never stall the GPU line."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload]
    # parse() raises warm-up to 3 for the arms proper; the in-line baseline is a short sample
    env = dict(os.environ, PTH_BENCH_CPU_LEG=f"{steps},{warmup}")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env).stdout.strip().splitlines()
        return json.loads(out[-1])["cpu_baseline"]
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port",
                "sample": f"CPU leg failed or exceeded {timeout} s: {type(e).__name__}"}


# ----------------------------------------------------------------------- GPU arm
def ncu_dram_bytes(path, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of `kernel` from a committed
    `ncu --csv --page raw` export; None if the file or the columns are missing."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        rows = list(csv.reader(open(path, newline="")))
        head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        names, units = rows[head], rows[head + 1]
        k, rd, wr = names.index("Kernel Name"), names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
        vals = [float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]]
                for r in rows[head + 2:] if len(r) > max(k, rd, wr) and kernel in r[k]]
        return sum(vals) / len(vals) if vals else None
    except Exception:  # noqa: BLE001
        return None


GAE_NCU_CSV = os.path.join("profiles", "gae_r02_raw.csv")


def gae_roofline(torch, ops, peaks):
    """GAE kernel at BASELINE configs[2]'s roofline size; inputs (2.7 GB) exceed the 126 MB L2."""
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
    traffic = ncu_dram_bytes(os.path.join(ROOT, GAE_NCU_CSV), "gae_tma_kernel")
    return {"kernel": "gae_tma_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": src, "shape": [T, N], "bytes_per_launch": nbytes,
            "ms_per_launch": ms, "traffic": traffic,
            "traffic_source": f"{GAE_NCU_CSV}: ncu dram__bytes_read.sum + dram__bytes_write.sum per launch at this shape"
                              if traffic else f"{GAE_NCU_CSV} not found"}


def ffma_peak(torch, _lib, device):
    """FP32 FFMA rate measured in this run (pth_debug_ffma_peak): the bound of the MLP kernels."""
    lib, ctx = _lib.load(), _lib.Context.get(device)
    ctas, iters = ctx.sm_count * 2 * 4, 1 << 15
    sink = torch.empty(ctas * 1024, dtype=torch.float32, device="cuda")
    best = None
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.pth_debug_ffma_peak(ctx.handle, sink.data_ptr(), ctas, iters, _lib.current_stream()),
                   "pth_debug_ffma_peak")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return 2.0 * ctas * 1024 * iters * 8 / (best / 1e3) / 1e12


def timed_iterations(torch, tr, steps, barrier, flush=None):
    """K iterations with CUDA events on the launching stream -> (ms total, phase means, decisions)."""
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks, decisions = [], 0
    for _ in range(steps):
        if flush is not None:
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
        decisions += tr.N * tr.T + (m_alt if tr.alt is not None else tr.N * tr.T)
    barrier()
    ms_total = sum(a.elapsed_time(d) for a, b, c, d in marks)
    ph = {"rollout_ms": sum(a.elapsed_time(b) for a, b, c, d in marks) / steps,
          "gae_ms": sum(b.elapsed_time(c) for a, b, c, d in marks) / steps,
          "update_ms": sum(c.elapsed_time(d) for a, b, c, d in marks) / steps}
    return ms_total, ph, decisions


def extra_config(torch, name, flush):
    """One of the other BASELINE configs on this GPU: 3 warm-up + 5 timed iterations."""
    from pantheonrl_b200.engine import PPOConfig, VecTrainer
    w = WORKLOADS[name]
    out = {"baseline": w["baseline"], "workload": w["name"], "unit": UNIT}
    sync = torch.cuda.synchronize
    if name == "rps_single":
        # the reference's own shape through the facade: PPO('MlpPolicy', env, seed=10).learn(...)
        from pantheonrl_b200.common.agents import OnPolicyAgent
        from pantheonrl_b200.envs import RPSEnv
        from pantheonrl_b200.ppo import PPO
        env = RPSEnv()
        env.add_partner_agent(OnPolicyAgent(PPO("MlpPolicy", env, seed=10)))
        ego = PPO("MlpPolicy", env, seed=10)
        ego.learn(total_timesteps=2048)
        sync()
        t0 = time.perf_counter()
        ego.learn(total_timesteps=3 * 2048)
        sync()
        dt = time.perf_counter() - t0
        out.update(value=2 * 3 * 2048 / dt, ms_per_step=dt / 3 * 1e3, steps=3, warmup=1,
                   api="PPO.learn (n_envs = 1): MultiAgentEnv.step -> OnPolicyAgent.get_action/update per decision, "
                       "one kernel call per forward, wall clock around learn()")
        return out
    kw = {"layout": "simple"} if w["env"] == "overcooked" else {}
    tr = VecTrainer(w["env"], w["n_envs"], PPOConfig(n_steps=w["n_steps"], n_epochs=N_EPOCHS, n_minibatches=N_MB),
                    seed=10, partner=w["partner"], **kw)
    for _ in range(3):
        tr.iteration()
    sync()
    ms, ph, dec = timed_iterations(torch, tr, 5, sync, flush)
    out.update(value=dec / (ms / 1e3), ms_per_step=ms / 5, steps=5, warmup=3, phases_ms=ph)
    del tr
    torch.cuda.empty_cache()
    return out


def variant_config(torch, kind, cpu=True):
    """ADAP.train / ModularAlgorithm.train at the reference's own shape (LiarsDice-v0, n_steps 2048, batch_size 64,
    10 epochs: 320 minibatches of 64 samples) on this GPU — one pth_ppo_update launch per train() (per partner phase
    for ModularAlgorithm, 2 partners) — next to the torch-CPU restatement of the same train() (oracle/sb3_torch.py,
    one epoch timed, all epochs are alike)."""
    import numpy as np
    from pantheonrl_b200 import _lib, update as up
    from pantheonrl_b200.spaces import MultiDiscrete, to_pth_space
    nvec, heads, M, BS, E, C, K, S, Pn = [7] * 6 + [7, 12] * 12, [7, 12], 2048, 64, 10, 3, 5, 32, 2
    sp = to_pth_space(MultiDiscrete(nvec), MultiDiscrete(heads))
    rs = np.random.RandomState(0)
    obs = np.zeros((M, 32), np.uint8)
    for i, n in enumerate(nvec):
        obs[:, i] = rs.randint(0, n, M)
    act = np.zeros((M, 4), np.uint8)
    act[:, 0], act[:, 1] = rs.randint(0, 7, M), rs.randint(0, 12, M)
    logp = (-np.log(84.0) + 0.1 * rs.randn(M)).astype(np.float32)
    adv, ret = rs.randn(M).astype(np.float32), rs.randn(M).astype(np.float32)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
    lib = _lib.load()
    n_mb = M // BS
    perm = up.perm_feistel(M, E, 10, 4)
    if kind == "adap":
        P = int(lib.pth_adap_param_count(sp, C))
        ws = up.UpdateWorkspace(sp, M, BS, context_size=C)
        st, dr = up.adap_draw(E * n_mb, K, C, "l2", 10, 0x20100, S=S, n_mb=n_mb, M=M, batch_size=BS)
        cx = rs.randn(M, C).astype(np.float32)
        extra = dict(loss_kind=_lib.PTH_LOSS_ADAP, context=dev(cx), context_loss_coeff=0.1, ctx_states=st, ctx_draws=dr)
        phases = 1
    else:
        P = int(lib.pth_modular_param_count(sp, Pn))
        ws = up.UpdateWorkspace(sp, M, BS, num_partners=Pn)
        extra = dict(loss_kind=_lib.PTH_LOSS_MODULAR, num_partners=Pn, marginal_reg_coef=0.5)
        phases = Pn
    params = dev((0.1 * rs.randn(P)).astype(np.float32))
    m, v = torch.zeros_like(params), torch.zeros_like(params)
    d = [dev(x) for x in (obs, act, logp, adv, ret)]

    def train():
        for q in range(phases):
            kw = dict(extra, partner_idx=q) if kind == "modular" else extra
            up.ppo_update(sp, params, m, v, 0, *d, perm, BS, ws, **kw)
    for _ in range(2):
        train()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        train()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out = {"workload": f"{'ADAP' if kind == 'adap' else 'ModularAlgorithm (2 partners)'}.train, LiarsDice-v0, n_steps 2048, "
                       "batch_size 64, 10 epochs (the reference's shape)",
           "ms_per_train": ms, "minibatch_steps_per_s": phases * E * n_mb / (ms / 1e3), "steps": 5, "warmup": 2}
    if cpu:
        from oracle import sb3_torch
        torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
        if kind == "adap":
            pol = sb3_torch.AdapMlpPolicy(nvec=nvec, heads=heads, context_size=C, seed=0)
            full = np.concatenate([obs[:, :30].astype(np.float32), cx], axis=1)
            t0 = time.perf_counter()
            sb3_torch.adap_train(pol, full, act[:, :2], logp, adv, ret, perm[:1].cpu().numpy(), BS, st.cpu().numpy(),
                                 dr.cpu().numpy(), context_loss_coeff=0.1)
        else:
            pol = sb3_torch.ModularMlpPolicy(nvec=nvec, heads=heads, num_partners=Pn, seed=0)
            buf = (obs[:, :30], act[:, :2], logp, adv, ret)
            t0 = time.perf_counter()
            sb3_torch.modular_train(pol, [buf] * Pn, [perm[:1].cpu().numpy()] * Pn, BS, marginal_reg_coef=0.5)
        dt = time.perf_counter() - t0
        out["cpu_port"] = {"ms_per_train": dt * E * 1e3, "sample": "1 of 10 epochs timed (32 minibatches per partner), x 10",
                           "cores": torch.get_num_threads(), "kind": "port"}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pantheonrl_b200 import _lib, ops
    from pantheonrl_b200.engine import PPOConfig
    from pantheonrl_b200.partner_set import make_trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # the contract is ONE stdout line: whatever NCCL logs (NCCL_DEBUG is the caller's choice) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass

    w = WORKLOADS[args.workload]
    if args.workload == "rps_single":
        raise SystemExit("rps_single is measured as an extra config of the default run (facade, N = 1)")
    n_envs = args.n_envs or w["n_envs"]
    cfg = PPOConfig(n_steps=w["n_steps"], n_epochs=N_EPOCHS, n_minibatches=N_MB)
    kw = {"layout": "simple"} if w["env"] == "overcooked" else {}
    partners = args.partners or world
    if partners % world:
        raise SystemExit("--partners must be a multiple of --gpus")
    if partners // world > 1:
        kw["partners_per_gpu"] = partners // world
    # envs sharded: rank r owns global envs [r*N, (r+1)*N) and its partner learner(s)
    tr = make_trainer(w["env"], n_envs, cfg, seed=10, partner=w["partner"], device=f"cuda:{local}",
                      env0=rank * n_envs, group=dist.group.WORLD if world > 1 else None,
                      exchange=args.exchange, ego_update=args.ego_update, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        tr.iteration()
    barrier()

    # ---- timed region: K iterations, device timing, phases on the launching stream
    launches0 = _lib.LAUNCHES
    # L2 flush between timed steps: 256 MB written (> the 126 MB L2), outside the per-step event brackets
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with ClockSampler(local) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        ms_total, ph, decisions = timed_iterations(torch, tr, args.steps, barrier, flush)
        t_wall = time.perf_counter() - t_wall0
    clocks = clk.summary()
    launches = _lib.LAUNCHES - launches0

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
    # every step uploads every learner's parameters + Adam moments from pinned host
    # memory, runs the iteration, downloads the updated state and the logged scalars.
    learners = tr.learners()
    host = [[t.cpu().pin_memory() for t in (L.params, L.adam_m, L.adam_v)] for L in learners]
    stats_host = torch.empty_like(tr.ego.last_stats, device="cpu").pin_memory()
    h2d = sum(t.numel() * 4 for v in host for t in v)
    d2h = h2d + stats_host.numel() * 4 * len(learners) + 16
    barrier()
    t0 = time.perf_counter()
    e2e_dec = 0
    for _ in range(args.steps):
        for L, hs in zip(learners, host):
            for dst, src in zip((L.params, L.adam_m, L.adam_v), hs):
                dst.copy_(src, non_blocking=True)
        e2e_dec += tr.iteration()
        for L, hs in zip(learners, host):
            for src, dst in zip((L.params, L.adam_m, L.adam_v), hs):
                dst.copy_(src, non_blocking=True)
        stats_host.copy_(tr.ego.last_stats, non_blocking=True)
        for L in learners[1:]:
            _ = L.last_stats.cpu()
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

    # ---- rooflines (rank 0, N = 1 only for the standalone GAE / FFMA measurements)
    roof = gae_roofline(torch, ops, peaks) if world == 1 else None
    flops = 3 * FWD_FLOPS[w["env"]] * N_EPOCHS * decisions / args.steps / world  # this GPU's learners, per step
    sm_mhz = clocks.get("sm_mhz") or 1500.0
    if world == 1:
        fp32_peak = ffma_peak(torch, _lib, local)
        peak_src = "measured in this run: pth_debug_ffma_peak (register-resident FFMA chains, all SMs, best of 6)"
    else:
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        peak_src = "148 SMs x 128 FFMA/clk x 2 x median SM clock under load (N > 1: not re-measured)"
    dom = {"kernel": "ppo_update_kernel (all learners of this GPU)",
           "bound": "fp32-ffma (not hbm/tensor: parity mode is fp32 CUDA-core math)",
           "achieved": flops / (ph["update_ms"] / 1e3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
           "frac": flops / (ph["update_ms"] / 1e3) / 1e12 / fp32_peak, "peak_source": peak_src,
           "flops_convention": "dense: 3 x forward FLOPs per sample per epoch (SURVEY.md 8d)"}

    cpu, extra = None, None
    if world == 1:
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_subprocess(args.workload)
        if not args.no_extra_configs and args.workload == "liar" and not args.n_envs and not args.partners:
            del tr
            torch.cuda.empty_cache()
            extra = []
            for name in ("rps_single", "rps_selfplay", "overcooked"):
                try:
                    e = extra_config(torch, name, flush)
                    if not args.no_cpu_baseline:
                        c = cpu_baseline_subprocess(name, steps=2, warmup=1, timeout=200)
                        e["cpu_port"] = {k: c.get(k) for k in ("value", "cores", "sample", "s_per_step")}
                except Exception as ex:  # noqa: BLE001
                    e = {"baseline": WORKLOADS[name]["baseline"], "error": f"{type(ex).__name__}: {ex}"}
                extra.append(e)
            for kind in ("adap", "modular"):  # the loss variants, at the reference's own shape
                try:
                    e = variant_config(torch, kind, cpu=not args.no_cpu_baseline)
                except Exception as ex:  # noqa: BLE001
                    e = {"error": f"{type(ex).__name__}: {ex}"}
                e["baseline"] = f"variant:{kind}"
                extra.append(e)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args, world),
        "phases_ms": ph, "wall_s": t_wall, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roof, "roofline_dominant": dom, "cpu_baseline": cpu, "configs": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        leg = os.environ.get("PTH_BENCH_CPU_LEG")  # in-line cpu_baseline of our arm: a short sample
        if leg:  # "steps,warmup[,envs]" (envs: the contract test's tiny sample)
            leg = [int(x) for x in leg.split(",")]
            a.steps, a.warmup = leg[:2]
            if len(leg) > 2:
                WORKLOADS[a.workload].update(cpu_envs=leg[2], cpu_mb=1)
        run_reference(a)
    else:
        run_ours(a)
