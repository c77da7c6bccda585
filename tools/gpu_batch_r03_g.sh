#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --warp-sampling-interval 1 --warp-sampling-buffer-size 536870912 --clock-control none --import-source on -k regex:"ffn3_kernel|cell3_kernel|conv_kernel" -c 4 -o gpurun_out/r03g_layer -f python tools/sanitize_run.py layer 32 1000 > gpurun_out/r03g_ncu.log 2>&1
tail -3 gpurun_out/r03g_ncu.log
