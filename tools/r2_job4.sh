#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/j4_pytest.log 2>&1
tail -4 gpurun_out/j4_pytest.log
python tools/ab_c2.py --rounds 2 \
  "r1:SDE_B200_RESIDENT_R1=1+SDE_UC=0,SDE_B200_DEFINES=SDE_UC=0" \
  "v2_nouc:SDE_B200_DEFINES=SDE_UC=0" \
  "v2_uc" \
  "v2_uc_q0:SDE_B200_DEFINES=SDE_RES_QUAD=0" \
  "f32_uc~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1" \
  "f32_uc_q0~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1+SDE_RES_QUAD=0" \
  "r1_uc:SDE_B200_RESIDENT_R1=1" \
  "f32_uc_b384~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1,block=384" \
  "f32_uc_b640~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1,block=640" \
  > gpurun_out/j4_ab.txt 2>&1
cat gpurun_out/j4_ab.txt
