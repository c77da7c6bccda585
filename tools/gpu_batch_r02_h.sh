#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest all gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --no-header 2>&1 | tail -12
echo "== tile golden verbose"; timeout 600 python -m pytest tests/test_tile_golden.py -m gpu -q -s --no-header -k "bf16_arm" 2>&1 | grep "^\[" 
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/r02h_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cell us', d['roofline']['us_per_call'], 'frac', d['roofline']['frac'], {k:(v['us']) for k,v in d['kernels'].items()}, d.get('roofline_step',{}).get('frac'))
print(json.dumps(d.get('other_configs'), indent=1))
PY
} > gpurun_out/r02h_main.log 2>&1
cat gpurun_out/r02h_main.log
