"""bf16 LayerNorm pass at 32 000 rows (D = 256, 512): time per call on rotating buffers (no L2 reuse)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S

dev = "cuda:0"
for D in (256, 512):
    ln = S.nnet.containers.LayerNorm(D).to(dev).eval()
    xs = [torch.randn(32000, D, device=dev).bfloat16() for _ in range(12)]
    with torch.no_grad():
        for i in range(12):
            ln(xs[i])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(12):
                ln(xs[i])
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 120 * 1e3
    print(f"D={D}: {us:.1f} us per call, {2 * 32000 * D * 2 / us / 1e6:.2f} TB/s")
