"""Timeline (clock64) of CTA 0 of K-CONV (conv_kernel) at the bench shape: per tile the events
0 tile start, 1 g staged, 2 depthwise + LN + act written (A operand), 3 all warps past the g tile, 4 accumulator + residual ready,
5 epilogue stored, 6 tile end."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S
from summarymixing_b200 import _lib as L

B, T, D = 32, 1000, 256
dev = "cuda:0"
torch.manual_seed(0)
m = S.ConvolutionModule(D, 31, True, S.Swish, 0.0, masked_false_or_true=False).eval().to(dev)
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
mask = torch.ones(B, T, 1, dtype=torch.bool, device=dev)
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3):
        m(x, mask)
    torch.cuda.synchronize()
    L.lib().smx_debug_set_trace(buf.data_ptr())
    m(x, mask)
    torch.cuda.synchronize()
    L.lib().smx_debug_set_trace(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m(x, mask)
    e1.record()
    torch.cuda.synchronize()
    print(f"conv module: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (same input every call: L2-warm)")
t = buf.cpu()[960:1024].view(4, 16)
nz = t[t > 0]
t0 = int(nz.min())
for it in range(4):
    ev = t[it]
    if int(ev.max()) == 0:
        continue
    v = [int(a) - t0 for a in ev[:7]]
    print(f"  tile {it}: " + " ".join(f"{a:7d}" for a in v) + "   deltas: " + " ".join(f"{b - a:6d}" for a, b in zip(v[:-1], v[1:])))
