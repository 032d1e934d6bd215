#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/j10_pytest.log 2>&1
grep -E "passed|failed|error|FAILED" gpurun_out/j10_pytest.log | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/j10_bench.json 2> gpurun_out/j10_bench.err
tail -3 gpurun_out/j10_bench.err
head -c 300 gpurun_out/j10_bench.json
