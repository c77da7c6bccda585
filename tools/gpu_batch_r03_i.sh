#!/bin/bash
tag=${1:-r03i}
mkdir -p gpurun_out
{
echo "== quick"; for args in "conv" "conv 32 1000" "conv 50 300" "layer 32 1000"; do timeout 180 python tools/sanitize_run.py $args 2>&1 | tail -1; done
echo "== pytest"; timeout 1500 python -m pytest tests/test_tc_path_gpu.py tests/test_tile_golden.py tests/test_fullsize_parity_gpu.py tests/test_tc_blocks_gpu.py -m gpu -q -x --no-header 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-others 2>&1 | tail -1 > gpurun_out/${tag}_bench.json; python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cell us', d['roofline']['us_per_call'], 'frac', d['roofline']['frac'], {k:(v['us']) for k,v in d['kernels'].items()})
PY
} > gpurun_out/${tag}_main.log 2>&1
cat gpurun_out/${tag}_main.log
