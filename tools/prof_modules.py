"""Call one module of the cfg2 encoder layer a few times at the bench shape (for ncu captures).
    python tools/prof_modules.py cell|ffn|conv|layer [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S

what = sys.argv[1] if len(sys.argv) > 1 else "cell"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, T, D = 32, 1000, 256
dev = "cuda:0"
torch.manual_seed(0)
layer = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval().to(dev)
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
lens = torch.randint(500, T + 1, (B,))
lens[0] = T
mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
with torch.no_grad():
    for _ in range(iters):
        if what == "cell":
            layer.mha_layer(x, src_padding_mask=mask)
        elif what == "conv":
            layer.convolution_module(x, mask.unsqueeze(-1))
        else:
            layer(x, src_key_padding_mask=mask)
torch.cuda.synchronize()
print("ok")
