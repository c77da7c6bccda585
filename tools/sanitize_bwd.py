"""Small training-path workload for compute-sanitizer (racecheck / memcheck): a Branchformer layer (conv-branch backward with the
reflect fix-up, dropout masks) and a Conformer layer (conv module backward: register-window depthwise kernels) forward + backward.

    compute-sanitizer --tool racecheck python tools/sanitize_bwd.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import summarymixing_b200 as S  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
B, T, D = 3, 140, 128
mask = (torch.arange(T)[None] < torch.tensor([T, 75, 33])[:, None]).to(dev)
for name, m in (("branchformer", S.BranchformerEncoderLayer(D, 2, 31, csgu_linear_units=256, local_proj_hid_dim=[D], local_proj_out_dim=D,
                                                            summary_hid_dim=[D], summary_out_dim=D, mode="SummaryMixing", dropout=0.1)),
                ("conformer", S.ConformerEncoderLayer(D, 256, 2, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                                      local_proj_out_dim=D, summary_hid_dim=[D], dropout=0.1))):
    m = m.to(dev).train()
    x = torch.randn(B, T, D, device=dev, requires_grad=True)
    y = m(x, src_key_padding_mask=mask)[0]
    y.pow(2).mean().backward()
    torch.cuda.synchronize()
    ok = all(p.grad is None or bool(torch.isfinite(p.grad).all()) for p in m.parameters()) and bool(torch.isfinite(x.grad).all())
    print(f"{name}: forward + backward done, gradients finite: {ok}")
