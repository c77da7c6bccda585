#!/bin/bash
mkdir -p gpurun_out
{
echo "== quick"; timeout 180 python tools/sanitize_run.py cell 50 300 2>&1 | tail -1
echo "== trace normal"; timeout 300 python tools/trace_cell4.py 2>&1 | tail -9
echo "== trace noweights"; SMX_DBG_C4_NOWEIGHTS=1 timeout 300 python tools/trace_cell4.py 2>&1 | tail -9
echo "== bench noweights"; SMX_DBG_C4_NOWEIGHTS=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-others 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cell us', d['roofline']['us_per_call'])"
} > gpurun_out/r02r_main.log 2>&1
cat gpurun_out/r02r_main.log
