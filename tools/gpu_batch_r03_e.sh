#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --warp-sampling-interval 1 --warp-sampling-buffer-size 536870912 --clock-control none --import-source on -k regex:conv_kernel -c 1 -o gpurun_out/r03e_conv -f python tools/sanitize_run.py conv 32 1000 > gpurun_out/r03e_ncu.log 2>&1
tail -3 gpurun_out/r03e_ncu.log
