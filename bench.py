"""bench.py — encoder-forward frames/s of the SummaryMixing-Conformer hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Placeholder of round 1, step 1: device-resident timing of the 12-layer D=256 encoder.  The full contract
(roofline, cpu_baseline, e2e, clocks, reference arm) is filled in as the kernels land.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smx")
    ap.add_argument("--dtype", default="fp32")
    args = ap.parse_args()
    import torch
    import summarymixing_b200 as S
    from summarymixing_b200 import _lib

    dev = "cuda:0"
    torch.manual_seed(0)
    B, T, D = 32, 1000, 256
    enc = S.ConformerEncoder(12, D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                             local_proj_out_dim=D, summary_hid_dim=[D], mode="SummaryMixing").eval().to(dev)
    g = torch.Generator().manual_seed(0)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    x = torch.randn(B, T, D, generator=g).to(dt).to(dev)
    lens = torch.randint(500, 1001, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
    with torch.no_grad():
        for _ in range(args.warmup):
            enc(x, src_key_padding_mask=mask)
        torch.cuda.synchronize()
        n0 = _lib.lib().smx_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            enc(x, src_key_padding_mask=mask)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"metric": "encoder-fwd frames/sec (B=32,T=1000,D=256)", "value": B * T / (ms / 1e3),
                      "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "higher_is_better": True, "dtype": args.dtype,
                      "gpu_launches": int(_lib.lib().smx_launch_count() - n0)}))


if __name__ == "__main__":
    main()
