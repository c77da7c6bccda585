#!/bin/bash
tag=${1:-r04a}
mkdir -p gpurun_out
{
echo "== ffn version sweep ring 128K"; SMX_F3_RING=131072 timeout 120 python tools/ffn_version_sweep.py 2>&1 | tail -4
echo "== ffn version sweep ring 64K"; SMX_F3_RING=65536 timeout 120 python tools/ffn_version_sweep.py 2>&1 | tail -4
echo "== ffn version sweep ring 96K"; SMX_F3_RING=98304 timeout 120 python tools/ffn_version_sweep.py 2>&1 | tail -4
} > gpurun_out/${tag}_main.log 2>&1
cat gpurun_out/${tag}_main.log
