"""Timeline (clock64) of CTA 0 of the persistent FFN kernel at the bench shape (first call in a layer = ffn1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S
from summarymixing_b200 import _lib as L

B, T, D = 32, 1000, 256
dev = "cuda:0"
torch.manual_seed(0)
layer = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval().to(dev)
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
mask = torch.ones(B, T, dtype=torch.bool, device=dev)
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3):
        layer(x, src_key_padding_mask=mask)
    torch.cuda.synchronize()
    if os.environ.get('SMX_FFN_VER'):
        L.lib().smx_debug_set_ffn_version(int(os.environ['SMX_FFN_VER']))
        layer(x, src_key_padding_mask=mask)
    L.lib().smx_debug_set_cell_version(3)  # (the one-kernel cell's own trace stamps would overlap the FFN's slots)
    L.lib().smx_debug_set_trace(buf.data_ptr())
    layer(x, src_key_padding_mask=mask)
    torch.cuda.synchronize()
    L.lib().smx_debug_set_trace(None)
t = buf.cpu()[512:512 + 5 * 2 * 32].view(5, 2, 32)   # the last FFN call of the layer (ffn2, with output LayerNorm) wins
roles = ["producer", "issuer", "prologue", "epi-g0", "epi-g1"]
nz = t[:, :, :16][t[:, :, :16] > 1000000]
t0 = int(nz.min())
for r in range(5):
    for it in range(2):
        ev = t[r, it]
        if int(ev.max()) == 0:
            continue
        print(f"  {roles[r]:9s} it{it}: " + " ".join(f"{int(v) - t0:7d}" if int(v) else "      -" for v in ev[:14]))

for it in range(2):
    w = t[1, it, 16:19]
    print(f"  issuer it{it} waited (cycles): weights {int(w[0])}, hidden chunk {int(w[1])}, drained accumulator {int(w[2])}")
for it in range(2):
    w = t[0, it, 16:18]
    print(f"  producer it{it}: blocked on empty slots {int(w[0])} of {int(w[1])} cycles")
fine = buf.cpu()[512 + 320:512 + 360]
print("issuer, chunk 2 of tile 0 (gemm1(3): a1e-wait b/a, 4 x (full-wait b/a); h_full wait b/a; gemm2(2): 2 x (full-wait b/a)):")
print("  " + " ".join(f"{int(v) - t0:7d}" if int(v) else "      -" for v in fine[:20]))

cv = buf.cpu()[960:1024].view(4, 16)
c0 = int(cv[cv > 0].min())
print("conv kernel (tile: start, g staged, conv done, g free, acc ready, epilogue math done, stored):")
for it in range(4):
    if int(cv[it].max()):
        print(f"  it{it}: " + " ".join(f"{int(v) - c0:7d}" if int(v) else "      -" for v in cv[it][:7]))
