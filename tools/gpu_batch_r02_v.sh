#!/bin/bash
mkdir -p gpurun_out
{
for f in 0 2 4 6; do
echo "== dbg $f"; SMX_DBG_C4_NOWEIGHTS=$f timeout 300 python tools/trace_cell4.py 2>&1 | grep -A3 "last chained call\|chained calls" | grep -v "^--" | tail -9
done
} > gpurun_out/r02v_main.log 2>&1
cat gpurun_out/r02v_main.log
