"""Cell forward + backward under a dynamic-chunk sum mask at the bench shape (B=32, T=1000, D=256, h=4): time per call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S

dev = "cuda:0"
torch.manual_seed(0)
B, T, D, chunk = 32, 1000, 256, 16
m = S.SummaryMixing(D, 4, [D], D, [D], D, activation=S.Swish, global_dropout=0.0).to(dev).train()
x = torch.randn(B, T, D, device=dev, requires_grad=True)
mask = torch.ones(B, T, dtype=torch.bool, device=dev)
ci = torch.arange(T) // chunk
for name, sm in (("no sum mask", None), ("chunk mask", (ci[None, :] <= ci[:, None]).float().to(dev)),
                 ("weights (products)", (torch.rand(T, T) + 0.1).to(dev))):
    for _ in range(2):
        m(x, sum_mask=sm, src_padding_mask=mask).pow(2).mean().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(x, sum_mask=sm, src_padding_mask=mask).pow(2).mean().backward()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:20s}: {e0.elapsed_time(e1) / 5:.2f} ms per forward + backward of the cell")
