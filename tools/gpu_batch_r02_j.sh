#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest all gpu"; timeout 1800 python -m pytest tests -m gpu -q -x --no-header 2>&1 | tail -8
echo "== fullsize verbose"; timeout 900 python -m pytest tests/test_fullsize_parity_gpu.py -m gpu -q -s --no-header 2>&1 | grep "^\[\|passed\|failed"
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/r02j_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cell us', d['roofline']['us_per_call'], 'frac', d['roofline']['frac'], {k:(v['us']) for k,v in d['kernels'].items()})
print(d.get('fp32_arm')); print({k:(v.get('ms_per_step'), v.get('frames_per_s')) for k,v in d.get('other_configs',{}).items()})
PY
} > gpurun_out/r02j_main.log 2>&1
cat gpurun_out/r02j_main.log
