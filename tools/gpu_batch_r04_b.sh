#!/bin/bash
# full GPU suite + smoke + default bench
tag=${1:-r04f}
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench"; (timeout 900 python bench.py) > gpurun_out/${tag}_bench_stdout.txt 2>&1; grep "^{" gpurun_out/${tag}_bench_stdout.txt | tail -1 > gpurun_out/${tag}_bench.json
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'cell us', d['roofline']['us_per_call'], 'frac', d['roofline']['frac'], {k:(v['us']) for k,v in d['kernels'].items()}, 'step frac', d['roofline_step']['frac'])
print(json.dumps(d.get('other_configs'))[:600])
print(json.dumps(d.get('train'))[:300])
PY
} > gpurun_out/${tag}_main.log 2>&1
cat gpurun_out/${tag}_main.log
