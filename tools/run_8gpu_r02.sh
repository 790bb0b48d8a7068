#!/bin/bash
# 8-GPU measurements of round 2 (gpurun --gpus 8): multi-GPU parity tests, weak-scaling bench at 8 / 4 / 2,
# phase timeline of the sharded ego update at world 8, the 8-partner configuration on 8 GPUs.
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8 or 2" > gpurun_out/r2_mgpu_pytest.log 2>&1; tail -5 gpurun_out/r2_mgpu_pytest.log
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_scale_$n.json 2> gpurun_out/r2_scale_$n.err
  tail -c 600 gpurun_out/r2_scale_$n.json; echo
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r2_scale_1.json 2> gpurun_out/r2_scale_1.err; tail -c 400 gpurun_out/r2_scale_1.json; echo
timeout 300 $TR --nproc-per-node 8 --master-port 29650 tools/prof_update_phases_mgpu.py > gpurun_out/r2_phases_mgpu8.txt 2>&1; tail -28 gpurun_out/r2_phases_mgpu8.txt
timeout 300 $TR --nproc-per-node 8 --master-port 29660 bench.py --gpus 8 --steps 10 --warmup 3 --exchange p2p > gpurun_out/r2_scale_8_p2p.json 2> gpurun_out/r2_scale_8_p2p.err; tail -c 400 gpurun_out/r2_scale_8_p2p.json; echo
