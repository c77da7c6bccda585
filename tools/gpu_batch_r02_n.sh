#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --sampling-interval 0 --clock-control none --import-source on -k regex:cell4_kernel -s 2 -c 1 -o gpurun_out/r02n_cell4 python tools/ncu_cell_capture.py > gpurun_out/r02n_ncu.log 2>&1
tail -2 gpurun_out/r02n_ncu.log
