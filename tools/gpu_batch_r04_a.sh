#!/bin/bash
tag=${1:-r04a}
mkdir -p gpurun_out
{
echo "== ffn version sweep"; timeout 120 python tools/ffn_version_sweep.py 2>&1 | tail -4
echo "== trace v4 (pairs)"; SMX_FFN_VER=4 timeout 120 python tools/trace_ffn.py 2>&1 | head -12
echo "== pytest ffn"; timeout 600 python -m pytest tests -x -q -m gpu -k "ffn or layer or encoder" 2>&1 | tail -5
} > gpurun_out/${tag}_main.log 2>&1
cat gpurun_out/${tag}_main.log
