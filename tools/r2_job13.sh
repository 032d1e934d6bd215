#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/j13_pytest.log 2>&1
grep -E "passed|failed|error|FAILED|Error" gpurun_out/j13_pytest.log | tail -12
python bench.py --steps 10 --warmup 3 > gpurun_out/j13_bench.json 2> gpurun_out/j13_bench.err
tail -3 gpurun_out/j13_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/j13_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['clocks'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms'), v.get('roofline'))
PY
