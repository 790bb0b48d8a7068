#!/bin/bash
# 8-GPU measurements, second pass of round 2 (gpurun --gpus 8): after the cross-rank statistic polls were
# parallelised.  Multi-GPU parity at world 8, Liar's Dice weak scaling 8 / 4 / 2 / 1, the other BASELINE
# workloads (RPS self-play, Overcooked) at 8 and 1 GPUs, BASELINE configs[4] (8 partners) over 8 / 4 / 2 GPUs,
# phase timeline of the sharded ego update at world 8.
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8" > gpurun_out/r2b_mgpu_pytest.log 2>&1; tail -3 gpurun_out/r2b_mgpu_pytest.log
port=29700
run() {  # name gpus args...
  name=$1; n=$2; shift 2; port=$((port + 1))
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 $B "$@" > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err
  else
    timeout 300 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n $B "$@" > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err
  fi
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2b_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"] / 1e6, 2), "M/s", round(d["ms_per_step"], 2), "ms", d.get("phases_ms"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for n in 8 4 2 1; do run liar_$n $n; done
for n in 8 1; do run rps_$n $n --workload rps_selfplay; done
for n in 8 1; do run oc_$n $n --workload overcooked; done
for n in 8 4 2; do run liar_p8_$n $n --partners 8; done
timeout 300 $TR --nproc-per-node 8 --master-port 29790 tools/prof_update_phases_mgpu.py > gpurun_out/r2b_phases_mgpu8.txt 2>&1; tail -27 gpurun_out/r2b_phases_mgpu8.txt
