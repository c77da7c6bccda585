"""Small tcgen05-arm workload for compute-sanitizer (racecheck / synccheck / memcheck); prints max-abs vs the oracle.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py [cell|layer|lite|fast]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import summarymixing_b200 as S  # noqa: E402
from oracle import smx_oracle as O  # noqa: E402
from oracle.seeded import fill_module, seeded_input  # noqa: E402
from summarymixing_b200 import _lib as L  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "cell"
if len(sys.argv) > 2:  # B T overrides (e.g. "cell 40 300": more tiles than SMs / two tiles per CTA)
    pass
dev = "cuda:0"
D, B, T = 256, 3, 300
if len(sys.argv) > 3:
    B, T = int(sys.argv[2]), int(sys.argv[3])
x = seeded_input(1, B, T, D)
lens = torch.tensor(([T, T // 2 + 20, 9] + [max(1, T - 7 * i) for i in range(B)])[:B])
mask = torch.arange(T)[None] < lens[:, None]
if what in ("conv", "ffn"):
    if what == "conv":
        m = S.ConvolutionModule(D, 31, True, S.Swish, 0.0, masked_false_or_true=False).eval()
        fill_module(m, 2)
        y_or = O.convolution_module(x, dict(m.state_dict()), "", act="swish", mask=mask.unsqueeze(-1))
        run = lambda mm, xx: mm(xx, mask.unsqueeze(-1).to(dev))
    else:  # the FFN half-step runs through a layer's entry point only: use the C ABI via a one-layer encoder's first module
        m = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                                    summary_hid_dim=[D]).eval()
        fill_module(m, 2)
        y_or = O.conformer_layer(x, dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
        run = lambda mm, xx: mm(xx, src_key_padding_mask=mask.to(dev))[0]
elif what == "layer":
    m = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                                summary_hid_dim=[D]).eval()
    fill_module(m, 2)
    y_or = O.conformer_layer(x, dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
    run = lambda mm, xx: mm(xx, src_key_padding_mask=mask.to(dev))[0]
else:
    mode = {"cell": "SummaryMixing", "lite": "SummaryMixing-lite", "fast": "SummaryMixing-fast"}[what]
    m = S.SummaryMixing(D, 4, [D], D, [D], D, activation=S.Swish, mode=mode, use_layernorm=(what != "fast")).eval()
    fill_module(m, 2)
    y_or = O.summary_mixing(x, dict(m.state_dict()), mode=mode, act="swish", src_padding_mask=mask, use_layernorm=(what != "fast"))
    run = lambda mm, xx: mm(xx, src_padding_mask=mask.to(dev)).contiguous()
m = m.to(dev)
n0 = L.lib().smx_tc_launch_count()
with torch.no_grad():
    y = run(m, x.to(torch.bfloat16).to(dev))
torch.cuda.synchronize()
print(f"{what}: {L.lib().smx_tc_launch_count() - n0} tcgen05 launches, max-abs vs oracle {float((y.float().cpu() - y_or).abs().max()):.3e}")
