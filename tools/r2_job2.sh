#!/bin/bash
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_blocks.py -m gpu -x -q -s 2>&1 | tail -15 ) > gpurun_out/j2_blocks.log 2>&1
cat gpurun_out/j2_blocks.log
./tools/ubench/dfma_operands > gpurun_out/j2_dfma_operands.txt 2>&1
cat gpurun_out/j2_dfma_operands.txt
python tools/ab_c2.py --rounds 2 \
  "r1:SDE_B200_RESIDENT_R1=1" \
  "v2q1:SDE_B200_DEFINES=SDE_RES_QUAD=1" \
  "v2q0:SDE_B200_DEFINES=SDE_RES_QUAD=0" \
  "v2q2:SDE_B200_DEFINES=SDE_RES_QUAD=2" \
  "f32q1~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1" \
  "f32q0~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1+SDE_RES_QUAD=0" \
  "f32q0lit~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1+SDE_RES_QUAD=0+SDE_KC_LITERAL=1" \
  "single~:icdf=single" \
  > gpurun_out/j2_ab.txt 2>&1
cat gpurun_out/j2_ab.txt
