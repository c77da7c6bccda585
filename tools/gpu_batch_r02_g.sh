#!/bin/bash
mkdir -p gpurun_out
{
echo "== trace"; timeout 300 python tools/trace_cell4.py 2>&1 | tail -8
echo "== pytest quick"; timeout 900 python -m pytest tests/test_tc_path_gpu.py tests/test_tile_golden.py -m gpu -q -x --no-header 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/r02g_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cell us', d['roofline']['us_per_call'], 'frac', d['roofline']['frac'], {k:(v['us']) for k,v in d['kernels'].items()}, d.get('roofline_step',{}).get('frac'))
PY
} > gpurun_out/r02g_main.log 2>&1
cat gpurun_out/r02g_main.log
