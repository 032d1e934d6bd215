#!/bin/bash
mkdir -p gpurun_out
for c in c2cp; do
ncu --set full --clock-control none --import-source on -k regex:sde_sim_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/j22_$c -f python tools/run_cfg.py $c 1 > gpurun_out/j22_ncu_$c.log 2>&1
ncu -i gpurun_out/j22_$c.ncu-rep --page raw --csv > gpurun_out/j22_${c}_raw.csv 2>/dev/null
ncu -i gpurun_out/j22_$c.ncu-rep --page source --csv > gpurun_out/j22_${c}_src.csv 2>/dev/null
tail -1 gpurun_out/j22_ncu_$c.log
done
