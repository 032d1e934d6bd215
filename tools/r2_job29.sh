#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/j29_pytest.log 2>&1
grep -E "passed|failed|error|FAILED|Error" gpurun_out/j29_pytest.log | tail -12
python bench.py --steps 20 --warmup 5 > gpurun_out/j29_bench.json 2> gpurun_out/j29_bench.err
tail -2 gpurun_out/j29_bench.err
python -c "
import json
d=json.load(open('gpurun_out/j29_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])
print(d['e2e']['value'], d['plan_create_ms'], d['timed_output_parity'])
for k,v in d['configs'].items(): print(k, v.get('value'), v.get('ms'), (v.get('roofline') or {}).get('frac'))
"
python bench.py --impl reference --steps 3 --warmup 1 | head -c 400
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
