#!/bin/bash
# sanitizers on the layer at HEAD (K-FFN on CTA pairs with the staged final epilogue), small and >148-tile shapes
mkdir -p gpurun_out
for what in "layer" "layer 50 300"; do
  tag=$(echo $what | tr ' ' '_')
  timeout 900 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_run.py $what > gpurun_out/r04_racecheck_${tag}.log 2>&1
  timeout 900 compute-sanitizer --tool synccheck --print-limit 6 python tools/sanitize_run.py $what > gpurun_out/r04_synccheck_${tag}.log 2>&1
  timeout 900 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_run.py $what > gpurun_out/r04_memcheck_${tag}.log 2>&1
done
tail -n 3 gpurun_out/r04_*check_*.log
