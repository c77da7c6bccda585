"""FFN kernel time at the bench shape for kernel versions 3 (single CTAs) and 4 (CTA pairs); outputs must be bit-identical."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import summarymixing_b200 as S
from summarymixing_b200 import _host as H
from summarymixing_b200 import _lib as L

B, T, D = 32, 1000, 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
layer = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval().to(dev)
x = torch.randn(B, T, D, device=dev).to(torch.bfloat16)
mask = torch.ones(B, T, dtype=torch.bool, device=dev)
with torch.no_grad():
    layer(x, src_key_padding_mask=mask)
lw = layer._wv.struct
lib = L.lib()
st = H.stream_ptr(dev)
ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
xs = [torch.randn(B, T, D, device=dev).to(torch.bfloat16) for _ in range(8)]
ref = None
for cl in (3, 4, 3, 4):
    lib.smx_debug_set_ffn_version(cl)
    y = torch.empty_like(x)
    for oln in (False, True):
        args = (C.byref(lw.ffn2), lw.act, L.BF16, B * T)
        def call(i, out):
            L.check(lib.smx_ffn_fwd(*args, xs[i % 8].data_ptr(), lw.norm2_w if oln else None, lw.norm2_b if oln else None, 1e-5,
                                    out.data_ptr(), ws.data_ptr(), ws.numel(), st))
        call(0, y)
        torch.cuda.synchronize()
        key = ("oln" if oln else "plain")
        if ref is None:
            ref = {}
        if key not in ref:
            ref[key] = y.clone()
        same = torch.equal(ref[key], y)
        outs = [torch.empty_like(x) for _ in range(8)]
        for i in range(3):
            call(i, outs[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            call(i, outs[i % 8])
        e1.record()
        torch.cuda.synchronize()
        print(f"version {cl} {key:5s}: {1e3 * e0.elapsed_time(e1) / 20:7.1f} us  identical-to-v3={same}")
