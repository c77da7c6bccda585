"""Host-side (cProfile) cost of a Branchformer training step: where the Python / ctypes time goes."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, T, D = 8, 1000, 512
enc = S.BranchformerEncoder(4, D, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[512], local_proj_out_dim=512,
                            summary_hid_dim=[512], summary_out_dim=512, mode="SummaryMixing-lite", dropout=0.1).to(dev).train()
opt = torch.optim.SGD(enc.parameters(), lr=0.02)
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
mask = torch.ones(B, T, dtype=torch.bool, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    y = enc(x, src_key_padding_mask=mask)[0]
    y.float().pow(2).mean().backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
