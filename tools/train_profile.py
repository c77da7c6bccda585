"""In-situ kernel time breakdown of one training step (torch.profiler / CUPTI: warm caches, real overlap), cfg2 encoder, bf16 activations.

    python tools/train_profile.py [layers] [conformer|branchformer]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import summarymixing_b200 as S

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 12
what = sys.argv[2] if len(sys.argv) > 2 else "conformer"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if what == "branchformer":   # the recipe's Branchformer-lite dims and dropout, B=8
    B, T, D = 8, 1000, 512
    enc = S.BranchformerEncoder(layers, D, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[512], local_proj_out_dim=512,
                                summary_hid_dim=[512], summary_out_dim=512, mode="SummaryMixing-lite", dropout=0.1).to(dev).train()
else:
    B, T, D = 32, 1000, 256
    enc = S.ConformerEncoder(layers, D, 4 * D, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                             summary_hid_dim=[D], dropout=0.0).to(dev).train()
opt = torch.optim.SGD(enc.parameters(), lr=0.02)
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
mask = (torch.arange(T, device=dev)[None] < torch.randint(T // 2, T + 1, (B,), device=dev)[:, None])
target = torch.randn(B, T, D, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    y = enc(x, src_key_padding_mask=mask)[0]
    loss = ((y.float() - target) * mask[..., None]).pow(2).mean()
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
print(f"{what}, {layers} layers, B={B} T={T} D={D}: {e0.elapsed_time(e1) / 3:.2f} ms per step (CUDA events, 3 steps)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages()]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"{'total us':>10} {'n':>5} {'mean us':>9} {'share':>6}  kernel   [{sum(r[1] for r in rows)} launches, {tot / 1e3:.1f} ms of kernel time]")
for k, n, t in rows[:40]:
    print(f"{t:10.1f} {n:5d} {t / n:9.1f} {100 * t / tot:5.1f}%  {k[:90]}")
