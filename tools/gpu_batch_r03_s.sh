#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --warp-sampling-interval 1 --warp-sampling-buffer-size 536870912 --clock-control none --import-source on -k regex:cell4_kernel -s 2 -c 1 -o gpurun_out/r03s_cell4 -f python tools/ncu_cell_capture.py > gpurun_out/r03s_ncu.log 2>&1
tail -2 gpurun_out/r03s_ncu.log
