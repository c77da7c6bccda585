#!/bin/bash
# K-GEMM: block-diagonal linears visit only their heads' K-blocks (cfg3 cells): parity + cfg3 / cfg4 throughput
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_tc_path_gpu.py tests/test_tile_golden.py tests/test_fullsize_parity_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -2
python -c "
import bench, torch, json
print(json.dumps(bench.other_configs(torch.device('cuda',0))))
" 2>&1 | tail -1
