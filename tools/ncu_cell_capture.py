"""Workload for the K-SM DRAM-traffic capture (run under ncu, see tools/gpu_batch_*.sh):

    flush  : read 512 MB (sum)            -- L2 holds nothing of x / y / the weights afterwards
    K-SM   : smx_mixing_block_fwd         -- norm1 + SummaryMixing cell + skip at the bench shape (cudaMemset + cell4_kernel)
    evict  : read 512 MB (sum)            -- pure read: its dram WRITE bytes are the dirty lines of y that were still in L2

The profiled window (cudaProfilerStart/Stop) holds exactly these launches; tools/ncu_traffic.py adds them up.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as BN
from summarymixing_b200 import _host as H
from summarymixing_b200 import _lib as L

dev = torch.device("cuda", 0)
B, T, D = BN.B, BN.T, BN.D
enc = BN.build_encoder().to(dev)
x, mask = BN.make_inputs(1000, 1)[0]
xb = x.to(torch.bfloat16).to(dev)
mk = mask.to(dev)
with torch.no_grad():
    enc(xb, src_key_padding_mask=mk)  # fills the weight structs / packed images
lib = L.lib()
lw = enc._wv.struct[0]
m8 = mk.to(torch.uint8).contiguous()
nb = lib.smx_mixing_block_workspace_bytes(C.byref(lw.cell), L.BF16, B, T, 0)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
y = torch.empty_like(xb)
big = torch.ones(128 * 1024 * 1024, dtype=torch.float32, device=dev)  # 512 MB
st = H.stream_ptr(dev)


def ksm():
    L.check(lib.smx_mixing_block_fwd(C.byref(lw.cell), lw.norm1_w, lw.norm1_b, L.BF16, B, T, xb.data_ptr(), m8.data_ptr(), None,
                                     y.data_ptr(), ws.data_ptr(), ws.numel(), st))


for _ in range(3):
    ksm()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
s0 = big.sum()
ksm()
s1 = big.sum()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", float(s0), float(s1))
