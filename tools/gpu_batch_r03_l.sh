#!/bin/bash
mkdir -p gpurun_out
for what in "conv" "layer"; do
  timeout 900 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_run.py $what > gpurun_out/r03k_racecheck_${what}.log 2>&1
done
tail -n 4 gpurun_out/r03k_racecheck_conv.log gpurun_out/r03k_racecheck_layer.log
timeout 600 python -m pytest tests/test_tc_path_gpu.py -m gpu -q -x --no-header 2>&1 | tail -2
