#!/bin/bash
mkdir -p gpurun_out
{
echo "== trace cta0"; timeout 300 python tools/trace_cell4.py 2>&1 | tail -8
echo "== trace cta120"; SMX_TRACE_CTA=120 timeout 300 python tools/trace_cell4.py 2>&1 | tail -6
} > gpurun_out/r02o_main.log 2>&1
cat gpurun_out/r02o_main.log
