#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/j18_pytest.log 2>&1
grep -E "passed|failed|error|FAILED|Error" gpurun_out/j18_pytest.log | tail -12
python tools/run_cfg.py c3 5 | tail -1
python tools/run_cfg.py c3t 5 | tail -1
python tools/run_cfg.py c2tpn 10 | tail -1
python tools/run_cfg.py c5p 3 | tail -1
