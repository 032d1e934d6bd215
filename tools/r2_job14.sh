#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/j14_pytest.log 2>&1
grep -E "passed|failed|error|FAILED|Error" gpurun_out/j14_pytest.log | tail -12
python tools/run_cfg.py c3 5
python tools/run_cfg.py c3t 5
ncu --set full --clock-control none --import-source on -k regex:sde_sim_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/j14_c3 -f python tools/run_cfg.py c3 1 > gpurun_out/j14_ncu.log 2>&1
ncu -i gpurun_out/j14_c3.ncu-rep --page raw --csv > gpurun_out/j14_c3_raw.csv 2>/dev/null
ncu -i gpurun_out/j14_c3.ncu-rep --page source --csv > gpurun_out/j14_c3_src.csv 2>/dev/null
tail -2 gpurun_out/j14_ncu.log
