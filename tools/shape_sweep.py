"""Shape sweep of the D = 256 Conformer layer on the bf16 tensor-core arm against the fp32-math arm of the same library (edge shapes:
fewer rows than a tile, ragged last tiles, more tiles than the resident kernels hold -> their fallbacks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S
from summarymixing_b200 import _lib as L

dev = "cuda:0"
D = 256
torch.manual_seed(5)
layer = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                                summary_hid_dim=[D]).eval().to(dev)
worst = 0.0
for B, T in [(1, 16), (1, 31), (2, 127), (1, 129), (3, 128), (5, 257), (37, 300), (300, 128), (40, 1000), (2, 5000), (148, 256), (149, 256)]:
    g = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, T, D, generator=g)
    lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
    with torch.no_grad():
        y32 = layer(x.to(dev), src_key_padding_mask=mask)[0]
        t0 = L.lib().smx_tc_launch_count()
        y16 = layer(x.to(torch.bfloat16).to(dev), src_key_padding_mask=mask)[0].float()
        n = L.lib().smx_tc_launch_count() - t0
    torch.cuda.synchronize()
    ref = layer(x.to(torch.bfloat16).float().to(dev), src_key_padding_mask=mask)[0] if False else y32
    valid = mask.unsqueeze(-1)
    err = float(((y16 - ref) * valid).abs().max())
    rel = float(((y16 - ref) * valid).norm() / (ref * valid).norm())
    worst = max(worst, rel)
    print(f"B={B:4d} T={T:5d}: {n} tcgen05 launches, max-abs {err:.3e} rel-L2 {rel:.3e}  finite={bool(torch.isfinite(y16).all())}")
    assert torch.isfinite(y16).all() and rel < 2e-2, (B, T, err, rel)
print("ok, worst rel-L2", worst)
