#!/bin/bash
# round 2, call A: new parity tests at real sizes + sanitizer logs of the current kernels
mkdir -p gpurun_out
python -m pytest tests/test_tile_golden.py tests/test_fullsize_parity_gpu.py -m gpu -q -s -x --no-header 2>&1 | tail -80 > gpurun_out/r02a_parity.log
for tool in racecheck synccheck memcheck; do
  for what in cell layer; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py $what > gpurun_out/r02a_sanitizer_${tool}_${what}.log 2>&1
    tail -3 gpurun_out/r02a_sanitizer_${tool}_${what}.log
  done
done
tail -40 gpurun_out/r02a_parity.log
