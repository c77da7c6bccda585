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
dev = "cuda:0"
D, B, T = 256, 3, 300
x = seeded_input(1, B, T, D)
lens = torch.tensor([300, 170, 9])
mask = torch.arange(T)[None] < lens[:, None]
if what == "layer":
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
