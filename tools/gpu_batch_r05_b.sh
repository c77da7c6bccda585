#!/bin/bash
# tile quantisation: two / four utterance lanes on concurrent streams against the single launch chain
mkdir -p gpurun_out
{
timeout 300 python tools/two_stream_check.py 1
SMX_C4_MAX_CTAS=64 timeout 300 python tools/two_stream_check.py 2
SMX_C4_MAX_CTAS=74 timeout 300 python tools/two_stream_check.py 2
SMX_C4_MAX_CTAS=37 timeout 300 python tools/two_stream_check.py 4
} > gpurun_out/r05b_lanes.log 2>&1
cat gpurun_out/r05b_lanes.log | tail -20
