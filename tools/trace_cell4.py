"""Timeline (clock64) of CTA 0 of the one-kernel cell (K-SM v4, smx_tc_cell4.cu) at the bench shape + module timing.

issuer events: one stamp before every half-GEMM it issues (phase 1: G1c0 G1c1 G2c0 G2c1 per tile; phase 2: + G3n0 G3n1).
epilogue events (first warp of each group; a group owns one chain): phase 1 per tile: start, after E1, after E2', after publish;
phase 2 per tile: start, after E1, after E2 (L stored), after c[b] fetched + row statistics, after E3.
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
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
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
slot0 = int((t[256:256 + 4 * 640].view(4, 640)[:, 0] > 0).nonzero()[0])


def cta_stamps(tt, slot):
    c = tt[256 + 640 * slot:256 + 640 * slot + 4 * 148].view(148, 4)
    return c[:, 0], c[:, 1], c[:, 2]


st, xr, en = cta_stamps(t, slot0)
ok = st > 0
if int(ok.sum()):
    s0 = int(st[ok].min())
    print(f"per-CTA wall clock (ns since the first CTA started): start max {int(st[ok].max()) - s0}, x tile landed median {int((xr[ok] - s0).median())} "
          f"max {int(xr[ok].max()) - s0}, end min {int(en[ok].min()) - s0} median {int((en[ok] - s0).median())} max {int(en[ok].max()) - s0}; "
          f"CTA 0: x {int(xr[0]) - s0} end {int(en[0]) - s0}; two-tile CTAs end median {int((en[:102] - s0).median())}, one-tile {int((en[102:148] - s0).median())}")
iss, epi = t[64:128], t[192:256]
print(f"traced CTA, single call: {int(t[61] - t[60])} cycles in {int(t[63] - t[62])} ns = {float(t[61] - t[60]) / max(1, int(t[63] - t[62])):.3f} GHz; "
      f"issuer waited {int(t[58])} cycles for weight steps, {int(t[59])} for epilogue hand-overs")
t[58:64] = 0
nz = t[:256][t[:256] > 0]
if nz.numel():
    t0 = int(nz.min())
    print("traced CTA:", os.environ.get("SMX_TRACE_CTA", "0"), " clock64 at first event", t0)
    print("producer (c0 wait x_dead / issue):", " ".join(f"{int(v) - t0:6d}" for v in t[0:4]))
    print("issuer after c_full:", " ".join(f"{int(v) - t0:6d}" for v in t[64 + 40:64 + 42]))
    print("issuer  :", " ".join(f"{int(v) - t0:6d}" for v in iss[:40] if int(v)))
    print("epilogue group 0:", " ".join(f"{int(v) - t0:6d}" for v in epi if int(v)))
    print("epilogue group 1:", " ".join(f"{int(v) - t0:6d}" for v in t[128:192] if int(v)))

# ---- four calls back to back (programmatic launch chain, rotating inputs), the way bench.py times the kernel ----
import ctypes as C

from summarymixing_b200 import _host as H

enc = S.ConformerEncoder(1, D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                         summary_hid_dim=[D], mode="SummaryMixing").eval().to(dev)
with torch.no_grad():
    enc(xs[0], src_key_padding_mask=mask)
lib = L.lib()
lw = enc._wv.struct[0]
m8 = mask.to(torch.uint8).contiguous()
xs12 = xs + [torch.randn(B, T, D, device=dev).to(torch.bfloat16) for _ in range(4)]  # 12 x 16.4 MB > L2
ys12 = [torch.empty_like(xs[0]) for _ in range(12)]
n = 12
nb = lib.smx_mixing_block_workspace_bytes(C.byref(lw.cell), L.BF16, B, T, 0)
ws = torch.empty(nb + n * (4096 + 8 * B), dtype=torch.uint8, device=dev)
xs_p = (C.c_void_p * n)(*[x.data_ptr() for x in xs12])
ys_p = (C.c_void_p * n)(*[y.data_ptr() for y in ys12])
stp = H.stream_ptr(torch.device(dev))
buf.zero_()
lib.smx_debug_set_trace(buf.data_ptr())
L.check(lib.smx_mixing_block_fwd_batch(C.byref(lw.cell), lw.norm1_w, lw.norm1_b, L.BF16, B, T, n, xs_p, m8.data_ptr(), ys_p, ws.data_ptr(), ws.numel(), stp))
torch.cuda.synchronize()
lib.smx_debug_set_trace(None)
t = buf.cpu()
rows = []
for slot in range(4):
    st, xr, en = cta_stamps(t, slot)
    if int((st > 0).sum()) == 148:
        rows.append((int(st.min()), int(st.median()), int(st.max()), int(xr.median()), int(xr.max()), int(en.min()), int(en.median()), int(en.max())))
rows.sort()
if rows:
    z = rows[0][0]
    print("chained calls (last four of twelve), ns since the first of them started: [first CTA start, median start, last start | x landed median, max | end min, median, max]")
    for r_ in rows:
        print("   ", " ".join(f"{v - z:7d}" for v in r_))
print(f"traced CTA, last chained call: {int(t[61] - t[60])} cycles in {int(t[63] - t[62])} ns = {float(t[61] - t[60]) / max(1, int(t[63] - t[62])):.3f} GHz; "
      f"issuer waited {int(t[58])} cycles for weight steps, {int(t[59])} for epilogue hand-overs")
t[58:64] = 0
nz = t[:256][t[:256] > 0]
if nz.numel():
    t0 = int(nz.min())
    print("last chained call, traced CTA (cycles):")
    print("  issuer  :", " ".join(f"{int(v) - t0:6d}" for v in t[64:104] if int(v)))
    print("  epilogue group 0:", " ".join(f"{int(v) - t0:6d}" for v in t[192:256] if int(v)))
    print("  epilogue group 1:", " ".join(f"{int(v) - t0:6d}" for v in t[128:192] if int(v)))
