#!/bin/bash
# full GPU test suite + sanitizers on the tensor-core arm at HEAD
mkdir -p gpurun_out
{
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q -x --no-header 2>&1 | tail -6
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
} > gpurun_out/r03k_main.log 2>&1
for what in "cell" "cell 50 300" "conv" "layer"; do
  tag=$(echo $what | tr ' ' '_')
  timeout 900 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_run.py $what > gpurun_out/r03k_racecheck_${tag}.log 2>&1
  timeout 900 compute-sanitizer --tool synccheck --print-limit 6 python tools/sanitize_run.py $what > gpurun_out/r03k_synccheck_${tag}.log 2>&1
  timeout 900 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_run.py $what > gpurun_out/r03k_memcheck_${tag}.log 2>&1
done
cat gpurun_out/r03k_main.log
tail -n 2 gpurun_out/r03k_*check_*.log
