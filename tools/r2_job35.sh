#!/bin/bash
for c in c3 c3t c2cp c5p c2tpn; do
  SDE_B200_NSTAGE=2 python tools/run_cfg.py $c 5 | tail -1
  python tools/run_cfg.py $c 5 | tail -1
done
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/j35_pytest.log 2>&1
grep -E "passed|failed|error|FAILED|Error" gpurun_out/j35_pytest.log | tail -8
