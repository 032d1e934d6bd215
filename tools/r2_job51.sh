#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sde_sim_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/j51_c2 -f python tools/run_cfg.py c2 1 > gpurun_out/j51_ncu.log 2>&1
ncu -i gpurun_out/j51_c2.ncu-rep --page raw --csv > gpurun_out/j51_c2_raw.csv 2>/dev/null
ncu -i gpurun_out/j51_c2.ncu-rep --page source --csv > gpurun_out/j51_c2_src.csv 2>/dev/null
tail -1 gpurun_out/j51_ncu.log
