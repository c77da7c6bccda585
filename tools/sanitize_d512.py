"""Small D = 512 workloads for compute-sanitizer: one conformer_large layer (K-GEMM incl. the GLU epilogue and block-diagonal K ranges,
LayerNorm pass, persistent tiled depthwise kernel) and one Branchformer-lite layer (K-GEMM with the GELU epilogue, CSGU gate kernel),
bf16, against the oracle.

    compute-sanitizer --tool racecheck python tools/sanitize_d512.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import summarymixing_b200 as S  # noqa: E402
from oracle import smx_oracle as O  # noqa: E402
from summarymixing_b200 import _lib as L  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
B, T, D = 3, 300, 512
x = torch.randn(B, T, D)
mask = torch.arange(T)[None] < torch.tensor([T, 170, 40])[:, None]
for name in ("conformer_large", "branchformer_lite"):
    if name == "conformer_large":
        m = S.ConformerEncoderLayer(D, 2048, 8, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                                    summary_hid_dim=[D]).eval()
        y_or = O.conformer_layer(x, dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
    else:
        m = S.BranchformerEncoderLayer(D, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[D], local_proj_out_dim=D, summary_hid_dim=[D],
                                       summary_out_dim=D, mode="SummaryMixing-lite").eval()
        with torch.no_grad():
            m.convolution_branch.csgu.conv.conv.weight.add_(0.1 * torch.randn(m.convolution_branch.csgu.conv.conv.weight.shape))
        y_or = O.branchformer_layer(x, dict(m.state_dict()), "", act="gelu", mode="SummaryMixing-lite", src_key_padding_mask=mask)
    m = m.to(dev)
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        y = m(x.to(torch.bfloat16).to(dev), src_key_padding_mask=mask.to(dev))[0]
    torch.cuda.synchronize()
    print(f"{name}: {L.lib().smx_tc_launch_count() - n0} tcgen05 launches, max-abs vs oracle {float((y.float().cpu() - y_or).abs().max()):.3e} "
          f"(|y|max {float(y_or.abs().max()):.2f})")
