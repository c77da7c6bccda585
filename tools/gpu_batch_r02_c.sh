#!/bin/bash
mkdir -p gpurun_out
{
echo "== tmem bench"; timeout 120 tools/micro/tmem_bench 2>&1
for args in "cell" "cell 32 1000" "cell 37 300"; do
  echo "== sanitize_run $args"; timeout 180 python tools/sanitize_run.py $args 2>&1 | tail -3
done
echo "== trace"; timeout 300 python tools/trace_cell4.py 2>&1 | tail -8
echo "== pytest tc tests"; timeout 900 python -m pytest tests/test_tc_path_gpu.py -m gpu -q -x --no-header 2>&1 | tail -8
} > gpurun_out/r02c_main.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_run.py conv > gpurun_out/r02c_racecheck_conv.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_run.py layer > gpurun_out/r02c_racecheck_layer.log 2>&1
cat gpurun_out/r02c_main.log
tail -n 3 gpurun_out/r02c_racecheck_conv.log gpurun_out/r02c_racecheck_layer.log
