#!/bin/bash
mkdir -p gpurun_out
for c in c5p c5; do
ncu --set full --clock-control none --import-source on -k regex:sde_sim_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/j56_$c -f python tools/run_cfg.py $c 1 > gpurun_out/j56_ncu_$c.log 2>&1
ncu -i gpurun_out/j56_$c.ncu-rep --page raw --csv > gpurun_out/j56_${c}_raw.csv 2>/dev/null
ncu -i gpurun_out/j56_$c.ncu-rep --page source --csv > gpurun_out/j56_${c}_src.csv 2>/dev/null
tail -1 gpurun_out/j56_ncu_$c.log
done
