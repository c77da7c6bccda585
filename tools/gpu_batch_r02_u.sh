#!/bin/bash
mkdir -p gpurun_out
{
echo "== trace cta 0"; timeout 300 python tools/trace_cell4.py 2>&1 | tail -14
echo "== trace cta 60"; SMX_TRACE_CTA=60 timeout 300 python tools/trace_cell4.py 2>&1 | tail -5
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
} > gpurun_out/r02u_main.log 2>&1
cat gpurun_out/r02u_main.log
