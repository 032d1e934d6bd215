#!/bin/bash
# round-2 GPU job 1: parity of the changed kernels, seed accuracy of the MUFU units, A/B of the C2 kernel variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/j1_gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/j1_pytest.log 2>&1
tail -5 gpurun_out/j1_pytest.log
./tools/ubench/mufu_seed > gpurun_out/j1_mufu_seed.txt 2>&1
cat gpurun_out/j1_mufu_seed.txt
python tools/ab_c2.py --rounds 2 \
  "r1:SDE_B200_RESIDENT_R1=1" \
  "v2" \
  "v2_donor:SDE_B200_DEFINES=SDE_SEED_DONOR_LOCAL=1" \
  "v2_lit:SDE_B200_DEFINES=SDE_KC_LITERAL=1" \
  "v2_b384:block=384" \
  "v2_b640:block=640" \
  "v2_nofold:SDE_B200_DEFINES=SDE_RES_FOLD=0" \
  > gpurun_out/j1_ab.txt 2>&1
cat gpurun_out/j1_ab.txt
