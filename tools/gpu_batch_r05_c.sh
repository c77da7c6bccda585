#!/bin/bash
# register-window depthwise kernels (fp32 arm: forward, data / weight gradients, reflect fix-up): parity + training step time
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dropout_gpu.py tests/test_backward_gpu.py tests/test_parity_gpu.py -x -q > gpurun_out/r05c_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r05c_tests.log
tail -n 6 gpurun_out/r05c_tests.log
timeout 600 python tools/train_step.py --steps 5 --warmup 2 --dtype bf16 > gpurun_out/r05c_train_bf16.log 2>&1
tail -n 2 gpurun_out/r05c_train_bf16.log
