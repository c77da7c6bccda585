#!/bin/bash
tag=${1:-r04g}
mkdir -p gpurun_out
{
echo "== pytest dropout"; timeout 900 python -m pytest tests/test_dropout_gpu.py tests/test_backward_gpu.py -x -q -m gpu 2>&1 | tail -25
} > gpurun_out/${tag}_main.log 2>&1
cat gpurun_out/${tag}_main.log
