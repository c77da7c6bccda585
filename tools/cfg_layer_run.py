"""One encoder of BASELINE configs[1] / [2] / [3] dims with a few layers, run a few times eagerly (for ncu launch lists: the
last 5 x launches-per-forward launches of the process are the five timed forwards):

    python tools/cfg_layer_run.py cfg2|cfg3|cfg4 [layers] [B] [T]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S

what = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
NL = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
T = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
dev = "cuda:0"
torch.manual_seed(3)
Dm = 256 if what == "cfg2" else 512
if what == "cfg2":
    enc = S.ConformerEncoder(NL, 256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256], local_proj_out_dim=256,
                             summary_hid_dim=[256], mode="SummaryMixing").eval().to(dev)
elif what == "cfg3":
    enc = S.ConformerEncoder(NL, 512, 2048, 8, 31, attention_type="SummaryMixing", local_proj_hid_dim=[512], local_proj_out_dim=512,
                             summary_hid_dim=[512], mode="SummaryMixing").eval().to(dev)
else:
    enc = S.BranchformerEncoder(NL, 512, 8, attention_type="SummaryMixing", csgu_linear_units=3072, local_proj_hid_dim=[512],
                                local_proj_out_dim=512, summary_hid_dim=[512], summary_out_dim=512, mode="SummaryMixing-lite").eval().to(dev)
x = torch.randn(B, T, Dm, device=dev).to(torch.bfloat16)
lens = torch.randint(T // 2, T + 1, (B,))
lens[0] = T
mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
with torch.no_grad():
    for _ in range(3):
        enc(x, src_key_padding_mask=mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        enc(x, src_key_padding_mask=mask)
    e1.record()
    torch.cuda.synchronize()
print(f"{what}: {NL} layer(s), B={B}, T={T}: {e0.elapsed_time(e1) / 5:.3f} ms per forward")
