"""BASELINE.json configs[4]: utterance-length sweep of the cfg2 encoder (12-layer SummaryMixing-Conformer, D=256), B=8,
T = 500 .. 8000 step 500, bf16, one B200.  Reports ms per forward and RTF = t / (B*T*0.04 s) (40 ms of audio per frame after
4x subsampling of 10 ms hops).  The reference's self-attention Conformer side of that config is not part of this repo."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
enc = bench.build_encoder().to(dev)
B = 8
rows = []
with torch.no_grad():
    for T in range(500, 8001, 500):
        g = torch.Generator().manual_seed(T)
        x = torch.randn(B, T, bench.D, generator=g).to(torch.bfloat16).to(dev)
        lens = torch.randint(T // 2, T + 1, (B,), generator=g)
        lens[0] = T
        mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
        for _ in range(3):
            enc(x, src_key_padding_mask=mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            enc(x, src_key_padding_mask=mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        rows.append({"T": T, "B": B, "ms": round(ms, 3), "frames_per_s": round(B * T / ms * 1e3), "rtf": ms / 1e3 / (B * T * 0.04)})
        print(f"T={T:5d}  {ms:8.3f} ms  {B * T / ms * 1e3 / 1e6:6.2f} M frames/s  RTF {rows[-1]['rtf']:.2e}")
print(json.dumps({"config": "cfg5 sweep (SummaryMixing side)", "rows": rows}))
