#!/bin/bash
# launch list of one training step (bf16 activations) at HEAD
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r05d_train_launches_all.csv python tools/train_step.py --steps 1 --warmup 2 --dtype bf16 > gpurun_out/r05d_train.log 2>&1
tail -1 gpurun_out/r05d_train.log | cut -c1-300
python tools/ncu_launch_summary.py gpurun_out/r05d_train_launches_all.csv 3750 > gpurun_out/r05d_train_launches_summary.txt
head -32 gpurun_out/r05d_train_launches_summary.txt
rm -f gpurun_out/r05d_train_launches_all.csv
