#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest all gpu"; timeout 2400 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -15
echo "== bwd bench"; timeout 600 python tools/bwd_bench.py 2>&1 | tail -6
echo "== train step"; timeout 600 python tools/train_step.py 2>&1 | tail -6
echo "== rtf sweep quick"; timeout 900 python tools/rtf_sweep.py --quick 2>&1 | tail -5
} > gpurun_out/r02k_main.log 2>&1
cat gpurun_out/r02k_main.log
