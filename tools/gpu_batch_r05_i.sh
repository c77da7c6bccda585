#!/bin/bash
# K-CONV with software-pipelined depthwise items: timeline, parity tests, bench
mkdir -p gpurun_out
python tools/trace_conv.py 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_tc_path_gpu.py tests/test_tile_golden.py tests/test_tc_blocks_gpu.py -x -q -m gpu 2>&1 | tail -2
python bench.py --no-train 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], {k:v['us'] for k,v in d['kernels'].items()})"
