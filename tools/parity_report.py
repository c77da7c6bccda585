"""Error of the bf16 tensor-core arm against the CPU oracle (fp32, same bf16-rounded input) at the bench shape: the numbers
quoted in DESIGN.md section 2.  The oracle is the checker here, exactly as in tests/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import summarymixing_b200 as S
from oracle import smx_oracle as O

DEV = "cuda:0"
torch.manual_seed(21)
D, B, T = 256, 32, 1000
g = torch.Generator().manual_seed(22)
x = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
lens = torch.randint(500, T + 1, (B,), generator=g); lens[0] = T
mask = torch.arange(T)[None] < lens[:, None]

def report(name, y, y_or):
    y = y.float().cpu()
    print(f"{name:42s} max-abs {float((y - y_or).abs().max()):.3e}  rel-L2 {float((y - y_or).norm() / y_or.norm()):.3e}  |y|max {float(y_or.abs().max()):.2f}")

with torch.no_grad():
    cell = S.SummaryMixing(D, 4, [D], D, [D], D, activation=S.Swish).eval()
    y_or = O.summary_mixing(x.float(), dict(cell.state_dict()), mode="SummaryMixing", act="swish", src_padding_mask=mask)
    report("SummaryMixing cell (K-SM)", cell.to(DEV)(x.to(DEV), src_padding_mask=mask.to(DEV)), y_or)
    layer = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D, summary_hid_dim=[D]).eval()
    y_or = O.conformer_layer(x.float(), dict(layer.state_dict()), "", act="swish", src_key_padding_mask=mask)
    report("Conformer layer (FFN+cell+conv+FFN+LN)", layer.to(DEV)(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0], y_or)
    y32 = layer(x.float().to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    report("  same layer, fp32-math arm", y32, y_or)
