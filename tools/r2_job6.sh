#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/j6_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/j6_pytest.log | tail -3
python tools/ab_c2.py --rounds 2 \
  "cubic_nouc_b512:SDE_B200_DEFINES=SDE_ICDF_F32SEED=0+SDE_UC=0,block=512" \
  "default~" \
  "horner~:SDE_B200_DEFINES=SDE_ICDF_HORNER=1" \
  "horner_b640~:SDE_B200_DEFINES=SDE_ICDF_HORNER=1,block=640" \
  "horner_b896~:SDE_B200_DEFINES=SDE_ICDF_HORNER=1,block=896" \
  > gpurun_out/j6_ab.txt 2>&1
cat gpurun_out/j6_ab.txt
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/j6_bench.json 2> gpurun_out/j6_bench.err
cat gpurun_out/j6_bench.json
