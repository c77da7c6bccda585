#!/bin/bash
# Branchformer training path (conv-branch backward, layer dropout) + regression of the backward / dropout suites
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dropout_gpu.py tests/test_backward_gpu.py -x -q > gpurun_out/r05a_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r05a_tests.log
tail -n 30 gpurun_out/r05a_tests.log
