"""bench.py's reference arm runs on the host cores (the oracle port of the reference's path), so its
JSON contract can be checked without a GPU: one line, the required keys, sane values."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(extra, env=None):
    # PTH_BENCH_CPU_LEG = "steps,warmup,envs": one step of an 8-env sample instead of the arm's real size
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"] + extra, capture_output=True, text=True,
                         timeout=600, env={**os.environ, "PTH_BENCH_CPU_LEG": "1,0,8", **(env or {})})
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_json_line():
    lines = run(["--gpus", "1"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec (all agents)" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("LiarsDice-v0 PPO-vs-PPO") and d["config"]["n_envs_per_gpu"] == 4096
    assert d["config"]["n_epochs"] == 10 and "all 10 epochs run" in d["cpu_baseline"]["sample"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "at 8 envs" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    assert run(["--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
