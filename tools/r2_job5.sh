#!/bin/bash
mkdir -p gpurun_out
python tools/ab_c2.py --rounds 2 \
  "uc_b512" \
  "uc_b768:block=768" \
  "uc_b1024:block=1024" \
  "f32_b512~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1" \
  "f32_b640~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1,block=640" \
  "f32_b768~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1,block=768" \
  "f32_b896~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1,block=896" \
  "f32_b1024~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1,block=1024" \
  "f32_b768_g8~:SDE_B200_DEFINES=SDE_ICDF_F32SEED=1,SDE_B200_RES_GRP=8,block=768" \
  > gpurun_out/j5_ab.txt 2>&1
cat gpurun_out/j5_ab.txt
