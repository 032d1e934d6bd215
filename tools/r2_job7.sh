#!/bin/bash
# 2 GPUs: full GPU test suite (multi-device tests included, NCCL path) + ncu captures of the C2 kernel
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/j7_gpus.txt
( time python -m pytest tests -m gpu -q ) > gpurun_out/j7_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/j7_pytest.log | tail -3
export CUDA_VISIBLE_DEVICES=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/j7_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/j7_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k sde_sim_kernel -s 1 -c 1 -o gpurun_out/r2_c2_sde_sim_kernel python tools/run_variant.py 16777216 0,0,0 > gpurun_out/j7_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
