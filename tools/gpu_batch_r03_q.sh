#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 4200 --csv --log-file gpurun_out/r03q_train_launches.csv python tools/train_step.py --steps 3 > gpurun_out/r03q_train.log 2>&1
tail -2 gpurun_out/r03q_train.log
python tools/ncu_launch_summary.py gpurun_out/r03q_train_launches.csv | head -40
