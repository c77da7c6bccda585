#!/bin/bash
# round 2, call B: first run of the one-kernel cell (K-SM v4)
mkdir -p gpurun_out
{
for args in "cell" "cell 40 300" "cell 32 1000" "cell 50 300" "layer 32 1000"; do
  echo "== sanitize_run $args"; timeout 180 python tools/sanitize_run.py $args 2>&1 | tail -3
done
echo "== trace"; timeout 300 python tools/trace_cell4.py 2>&1 | tail -8
echo "== pytest cell tests"; timeout 900 python -m pytest tests/test_tc_path_gpu.py -m gpu -q -x --no-header 2>&1 | tail -15
} > gpurun_out/r02b_main.log 2>&1
for what in "cell" "cell 50 300"; do
  tag=$(echo $what | tr ' ' '_')
  timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_run.py $what > gpurun_out/r02b_racecheck_${tag}.log 2>&1
  timeout 600 compute-sanitizer --tool synccheck --print-limit 10 python tools/sanitize_run.py $what > gpurun_out/r02b_synccheck_${tag}.log 2>&1
  timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_run.py $what > gpurun_out/r02b_memcheck_${tag}.log 2>&1
done
for what in conv; do
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 6 python tools/sanitize_run.py $what > gpurun_out/r02b_racecheck_${what}.log 2>&1
done
cat gpurun_out/r02b_main.log
tail -2 gpurun_out/r02b_*check_*.log
