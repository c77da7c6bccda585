"""Timeline (clock64) of CTA 0 of the GLU pass (cell3_kernel<2>) and the conv kernel at the bench shape."""
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
t = buf.cpu()[:512].view(8, 4, 16)[:5]
roles = ["producer", "issuer", "prologue", "epilogue", "-"]
nz = t[t > 0]
t0 = int(nz.min())
print("GLU pass: cycles since first event (issuer: GEMM start/end; epilogue: tile start, staged+LN, acc ready, stored)")
for r in range(5):
    for it in range(4):
        ev = t[r, it]
        if int(ev.max()) == 0:
            continue
        print(f"  {roles[r]:9s} it{it}: " + " ".join(f"{int(v) - t0:7d}" if int(v) else "      -" for v in ev[:13]))
