#!/bin/bash
# Dynamic Chunk Training backward (sum_mask cell, chunked convolution): gradient goldens + regression of the backward / dropout suites
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_backward_gpu.py tests/test_dropout_gpu.py tests/test_parity_gpu.py -x -q > gpurun_out/r05g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r05g_tests.log
grep -E "passed|failed|Error|rc=" gpurun_out/r05g_tests.log | tail -8
