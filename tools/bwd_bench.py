"""Times the SummaryMixing cell backward (smx_summary_mixing_bwd through the module surface) at the BASELINE cell shape.

    python tools/bwd_bench.py [--B 32 --T 1000 --D 256 --heads 4 --iters 10]

CUDA events on the current stream, warm, per call; prints one JSON line per I/O dtype.  The backward is the first-correct
fp32-math arm (generic strided GEMM): this tool exists to put a number next to it, not to claim a roofline."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import summarymixing_b200 as S  # noqa: E402
from summarymixing_b200 import _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--D", type=int, default=256)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = S.SummaryMixing(a.D, a.heads, [a.D], a.D, [a.D], a.D, activation=S.Swish).to(dev).eval()
    lens = torch.randint(a.T // 2, a.T + 1, (a.B,))
    mask = (torch.arange(a.T)[None] < lens[:, None]).to(dev)
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(a.B, a.T, a.D, device=dev).to(dt).requires_grad_(True)
        dy = torch.randn(a.B, a.T, a.D, device=dev).to(dt)
        with torch.no_grad():
            for _ in range(3):
                m(x.detach(), src_padding_mask=mask)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_f = t_b = 0.0
        n0 = L.lib().smx_launch_count()
        for i in range(a.iters + 2):
            x.grad = None
            ev[0].record()
            y = m(x, src_padding_mask=mask)
            ev[1].record()
            y.backward(dy)
            ev[2].record()
            torch.cuda.synchronize()
            if i >= 2:
                t_f += ev[0].elapsed_time(ev[1])
                t_b += ev[1].elapsed_time(ev[2])
        launches = (L.lib().smx_launch_count() - n0) // (a.iters + 2)
        frames = a.B * a.T
        print(json.dumps({"what": "SummaryMixing cell fwd+bwd", "io": str(dt).split(".")[-1], "B": a.B, "T": a.T, "D": a.D,
                          "heads": a.heads, "fwd_ms": t_f / a.iters, "bwd_ms": t_b / a.iters,
                          "frames_per_s_fwd_bwd": frames / ((t_f + t_b) / a.iters * 1e-3), "launches_per_step": int(launches),
                          "bwd_workspace_MB": L.lib().smx_summary_mixing_bwd_workspace_bytes(m._weights(dev), 0, a.B, a.T) / 1e6}))


if __name__ == "__main__":
    main()
