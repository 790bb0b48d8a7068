#!/bin/bash
# ncu evidence of round 2 (run under gpurun on ONE GPU).  Reports land in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/launches_r02.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:ppo_update -s 2 -c 1 -o gpurun_out/upd_r02 -f \
    python tools/prof_iter.py liar 4096 128 1 > gpurun_out/upd_r02.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/roll_r02 -f \
    python tools/prof_iter.py liar 4096 128 1 > gpurun_out/roll_r02.out 2>&1
ncu --set full --clock-control none -k regex:gae_tma -c 1 -o gpurun_out/gae_r02 -f \
    python tools/prof_gae.py > gpurun_out/gae_r02.out 2>&1
ncu --set full --clock-control none -k 'regex:policy_forward|pack_kernel|gae_ragged' -s 2 -c 12 -o gpurun_out/misc_r02 -f \
    python tools/prof_misc.py > gpurun_out/misc_r02.out 2>&1
ncu --set full --clock-control none -k 'regex:rollout_overcooked|ppo_update' -s 2 -c 2 -o gpurun_out/oc_r02 -f \
    python tools/prof_iter.py overcooked 1024 400 1 > gpurun_out/oc_r02.out 2>&1
ls -la gpurun_out/*.ncu-rep
# keep what is read offline small: text / CSV exports, then drop the big reports (gpurun_out is capped at 64 MiB)
for r in upd_r02 roll_r02 gae_r02 misc_r02 oc_r02; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
python tools/ncu_hot.py gpurun_out/upd_r02.ncu-rep ppo_update 70 > gpurun_out/upd_r02_hot.txt 2>&1
python tools/ncu_hot.py gpurun_out/roll_r02.ncu-rep rollout_kernel 40 > gpurun_out/roll_r02_hot.txt 2>&1
python tools/ncu_inst.py gpurun_out/upd_r02.ncu-rep ppo_update > gpurun_out/upd_r02_inst.txt 2>&1
rm -f gpurun_out/roll_r02.ncu-rep gpurun_out/gae_r02.ncu-rep gpurun_out/misc_r02.ncu-rep gpurun_out/oc_r02.ncu-rep
ls -la gpurun_out | head -40
