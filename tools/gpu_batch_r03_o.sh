#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest cfg3/cfg4/fullsize + tc tests"; timeout 2400 python -m pytest tests/test_tc_path_gpu.py tests/test_tile_golden.py tests/test_fullsize_parity_gpu.py tests/test_tc_blocks_gpu.py tests/test_parity_gpu.py -m gpu -q -x --no-header -s 2>&1 | grep "^\[cfg\|^\[conformer_layer_d512\|passed\|failed\|Error\|assert" | head -20
for c in cfg3 cfg4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r03o_${c}_launches.csv python tools/cfg_layer_run.py $c 1 > gpurun_out/r03o_${c}.log 2>&1
tail -1 gpurun_out/r03o_${c}.log
python tools/ncu_launch_summary.py gpurun_out/r03o_${c}_launches.csv | head -14
timeout 300 python tools/cfg_layer_run.py $c 1 | tail -1
done
} > gpurun_out/r03o_main.log 2>&1
cat gpurun_out/r03o_main.log
