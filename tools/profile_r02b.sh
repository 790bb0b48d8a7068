#!/bin/bash
# ncu evidence, second pass of round 2 (one GPU): the update kernel after the reduce / weight-load changes, and the
# ADAP / ModularAlgorithm variants of the kernel at the reference's shape.  Reports land in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
ncu --set full --clock-control none --import-source on -k regex:ppo_update -s 2 -c 1 -o gpurun_out/upd_r02b -f \
    python tools/prof_iter.py liar 4096 128 1 > gpurun_out/upd_r02b.out 2>&1
ncu --set full --clock-control none -k regex:ppo_update -s 4 -c 1 -o gpurun_out/adap_r02 -f \
    python tools/prof_variants.py > gpurun_out/adap_r02.out 2>&1
ncu --set full --clock-control none -k regex:ppo_update -s 18 -c 1 -o gpurun_out/mod_r02 -f \
    python tools/prof_variants.py > gpurun_out/mod_r02.out 2>&1
for r in upd_r02b adap_r02 mod_r02; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
python tools/ncu_hot.py gpurun_out/upd_r02b.ncu-rep ppo_update 50 > gpurun_out/upd_r02b_hot.txt 2>&1
rm -f gpurun_out/adap_r02.ncu-rep gpurun_out/mod_r02.ncu-rep
ls -la gpurun_out | grep r02
