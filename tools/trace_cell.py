"""Timeline (clock64) of CTA 0 of the fused cell kernels at the bench shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S
from summarymixing_b200 import _lib as L

B, T, D = 32, 1000, 256
dev = "cuda:0"
torch.manual_seed(0)
m = S.SummaryMixing(D, 4, [D], D, [D], D, activation=S.Swish).eval().to(dev)
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
lens = torch.randint(500, T + 1, (B,))
lens[0] = T
mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3):
        m(x, src_padding_mask=mask)
    torch.cuda.synchronize()
    L.lib().smx_debug_set_trace(buf.data_ptr())
    m(x, src_padding_mask=mask)
    torch.cuda.synchronize()
    L.lib().smx_debug_set_trace(None)
t = buf.cpu().view(2, 8, 4, 16)[:, :5]
roles = ["producer", "issuer", "prologue", "epilogue", "epi-g1"]
for ph in range(2):
    nz = t[ph][t[ph] > 0]
    if nz.numel() == 0:
        continue
    t0 = int(nz.min())
    print(f"pass {'AB'[ph]}: cycles since first event")
    for r in range(5):
        for it in range(4):
            ev = t[ph, r, it]
            if int(ev.max()) == 0:
                continue
            print(f"  {roles[r]:9s} it{it}: " + " ".join(f"{int(v) - t0:7d}" if int(v) else "      -" for v in ev[:10]))

sys.exit(0)
raw = buf.cpu().view(2, 512)
for ph in range(2):
    st = raw[ph, 320:320 + 192].view(48, 4)
    nz = t[ph][t[ph] > 0]
    t0 = int(nz.min())
    print(f"pass {'AB'[ph]} issuer steps (tile 0): before-full-wait, after-wait, after-mma, after-commit")
    for i in range(48):
        if int(st[i].max()) == 0:
            break
        print(f"  step {i:2d}: " + " ".join(f"{int(v) - t0:7d}" for v in st[i]))
