#!/bin/bash
# new training-path tests + sanitizers on the new backward kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py -x -q -k "branchformer" > gpurun_out/r05f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r05f_tests.log; tail -n 12 gpurun_out/r05f_tests.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_bwd.py > gpurun_out/r05_racecheck_bwd.log 2>&1; tail -n 4 gpurun_out/r05_racecheck_bwd.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_bwd.py > gpurun_out/r05_memcheck_bwd.log 2>&1; tail -n 4 gpurun_out/r05_memcheck_bwd.log
