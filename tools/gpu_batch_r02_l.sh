#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cell4_kernel -s 2 -c 1 -o gpurun_out/r02l_cell4 python tools/ncu_cell_capture.py > gpurun_out/r02l_ncu.log 2>&1
tail -3 gpurun_out/r02l_ncu.log
timeout 300 python tools/trace_cell4.py 2>&1 | tail -6
