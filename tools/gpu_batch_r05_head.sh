#!/bin/bash
# evidence at HEAD of the last session: full GPU test suite, smoke, bench (default + reference arm)
tag=${1:-r05_head}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; (time timeout 2400 python -m pytest tests -m gpu -q) > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench (default)"; (time python bench.py) > gpurun_out/${tag}_bench_stdout.txt 2>&1; grep "^{" gpurun_out/${tag}_bench_stdout.txt | tail -1 > gpurun_out/${tag}_bench.json; tail -4 gpurun_out/${tag}_bench_stdout.txt | grep real
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cell us', d['roofline']['us_per_call'], 'frac', d['roofline']['frac'], {k:(v['us']) for k,v in d['kernels'].items()}, 'step frac', d['roofline_step']['frac'])
print(json.dumps(d.get('other_configs')))
print(json.dumps(d.get('train'))[:900])
PY
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/${tag}_bench_reference_arm.json; cut -c1-300 gpurun_out/${tag}_bench_reference_arm.json
} > gpurun_out/${tag}_main.log 2>&1
cat gpurun_out/${tag}_main.log
