"""Timeline (clock64) of CTA 0 of the one-kernel cell (K-SM v4, smx_tc_cell4.cu) at the bench shape + module timing.

issuer events: one stamp before every half-GEMM it issues (phase 1: G1c0 G1c1 G2c0 G2c1 per tile; phase 2: + G3n0 G3n1).
epilogue events (warp 0): phase 1 per tile: start, after E1 x2, after E2' x2 + publish; phase 2 per tile: start, after E1 x2,
after E2 (L stored), after c[b] fetched, after E3 x2.
"""
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
xs = [torch.randn(B, T, D, device=dev).to(torch.bfloat16) for _ in range(8)]  # 131 MB: rotates through L2
lens = torch.randint(500, T + 1, (B,))
lens[0] = T
mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
buf = torch.zeros(2048, dtype=torch.int64, device=dev)
for ver in (4, 3):
    L.lib().smx_debug_set_cell_version(ver)
    with torch.no_grad():
        for i in range(5):
            m(xs[i % 8], src_padding_mask=mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40):
            m(xs[i % 8], src_padding_mask=mask)
        e1.record()
        torch.cuda.synchronize()
        print(f"cell version {ver}: {e0.elapsed_time(e1) / 40 * 1e3:.1f} us per call (eager launches, inputs rotating)")
L.lib().smx_debug_set_cell_version(4)
with torch.no_grad():
    L.lib().smx_debug_set_trace(buf.data_ptr())
    m(xs[0], src_padding_mask=mask)
    torch.cuda.synchronize()
    L.lib().smx_debug_set_trace(None)
t = buf.cpu()
cta = t[256:256 + 4 * 148].view(148, 4)
st, xr, en = cta[:, 0], cta[:, 1], cta[:, 2]
ok = st > 0
if int(ok.sum()):
    s0 = int(st[ok].min())
    print(f"per-CTA wall clock (ns since the first CTA started): start max {int(st[ok].max()) - s0}, x tile landed median {int((xr[ok] - s0).median())} "
          f"max {int(xr[ok].max()) - s0}, end min {int(en[ok].min()) - s0} median {int((en[ok] - s0).median())} max {int(en[ok].max()) - s0}; "
          f"CTA 0: x {int(xr[0]) - s0} end {int(en[0]) - s0}; two-tile CTAs end median {int((en[:102] - s0).median())}, one-tile {int((en[102:148] - s0).median())}")
iss, epi = t[64:128], t[192:256]
nz = t[t > 0]
if nz.numel():
    t0 = int(nz.min())
    print("traced CTA:", os.environ.get("SMX_TRACE_CTA", "0"), " clock64 at first event", t0)
    print("producer (c0 wait x_dead / issue):", " ".join(f"{int(v) - t0:6d}" for v in t[0:4]))
    print("issuer after c_full:", " ".join(f"{int(v) - t0:6d}" for v in t[64 + 40:64 + 42]))
    print("issuer  :", " ".join(f"{int(v) - t0:6d}" for v in iss[:40] if int(v)))
    print("epilogue:", " ".join(f"{int(v) - t0:6d}" for v in epi if int(v)))
