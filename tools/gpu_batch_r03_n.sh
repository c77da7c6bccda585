#!/bin/bash
mkdir -p gpurun_out
for c in cfg3 cfg4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r03n_${c}_launches.csv python tools/cfg_layer_run.py $c 1 > gpurun_out/r03n_${c}.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r03n_${c}_launches.csv | head -16
done
