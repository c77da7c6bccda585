"""BASELINE.json configs[4]: utterance-length sweep, SummaryMixing vs the reference's self-attention Conformer, one B200.

Both arms: 12-layer Conformer encoder, D=256, d_ffn=1024, h=4, k=31, B=8, T = 500 .. 8000 step 500, bf16 activations.
  * summarymixing : this repo's encoder (libsmx kernels through summarymixing_b200.ConformerEncoder)
  * mhsa          : the reference's ConformerEncoder(attention_type="regularMHA") algorithm (Conformer.py:425-429, 528-541) as
                    its own PyTorch path would run on the GPU: the oracle restatement (oracle.conformer_encoder_mhsa, pinned to the
                    unmodified reference by tests/golden/mhsa/) executed with torch CUDA ops -- cuBLAS linears, torch SDPA
                    (flash attention), cuDNN depthwise conv.  Library kernels, the comparison baseline, not product code.
Both arms are timed as CUDA-graph replays (and, for reference, with eager launches).  Reports ms per forward, frames/s and RTF = t / (B*T*0.04 s) (40 ms of audio per frame after 4x subsampling of 10 ms hops).

    python tools/rtf_sweep.py [--quick]     # --quick: T = 1000, 4000, 8000 only
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import summarymixing_b200 as S
from oracle import smx_oracle as O  # comparison arm only

dev = torch.device("cuda", 0)
B, D, NL, H = 8, bench.D, bench.LAYERS, bench.HEADS
enc = bench.build_encoder().to(dev)

# self-attention arm: state_dict with the reference's key names / shapes, random values, bf16 on the GPU
torch.manual_seed(1)
F, k = bench.FFN, bench.KSIZE
sd = {}
for i in range(NL):
    p = f"layers.{i}."
    shapes = {"mha_layer.att.in_proj_weight": (3 * D, D), "mha_layer.att.in_proj_bias": (3 * D,), "mha_layer.att.out_proj.weight": (D, D),
              "mha_layer.att.out_proj.bias": (D,), "convolution_module.layer_norm.weight": (D,), "convolution_module.layer_norm.bias": (D,),
              "convolution_module.bottleneck.0.weight": (2 * D, D, 1), "convolution_module.bottleneck.0.bias": (2 * D,),
              "convolution_module.conv.weight": (D, 1, k), "convolution_module.conv.bias": (D,),
              "convolution_module.after_conv.0.weight": (D,), "convolution_module.after_conv.0.bias": (D,),
              "convolution_module.after_conv.2.weight": (D, D), "convolution_module.after_conv.2.bias": (D,),
              "norm1.norm.weight": (D,), "norm1.norm.bias": (D,), "norm2.norm.weight": (D,), "norm2.norm.bias": (D,)}
    for m in ("ffn_module1", "ffn_module2"):
        shapes.update({f"{m}.0.weight": (D,), f"{m}.0.bias": (D,), f"{m}.1.ffn.0.weight": (F, D), f"{m}.1.ffn.0.bias": (F,),
                       f"{m}.1.ffn.3.weight": (D, F), f"{m}.1.ffn.3.bias": (D,)})
    for kk, shp in shapes.items():
        t = torch.randn(shp) * (1.0 / (shp[1] * (shp[2] if len(shp) > 2 else 1)) ** 0.5 if len(shp) > 1 else 0.1)
        if len(shp) == 1 and kk.endswith("weight"):
            t = 1.0 + t
        sd[p + kk] = t.to(torch.bfloat16).to(dev)
sd["norm.norm.weight"] = torch.ones(D, dtype=torch.bfloat16, device=dev)
sd["norm.norm.bias"] = torch.zeros(D, dtype=torch.bfloat16, device=dev)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


Ts = [1000, 4000, 8000] if "--quick" in sys.argv else list(range(500, 8001, 500))
rows = []
with torch.no_grad():
    for T in Ts:
        g = torch.Generator().manual_seed(T)
        x = torch.randn(B, T, D, generator=g).to(torch.bfloat16).to(dev)
        lens = torch.randint(T // 2, T + 1, (B,), generator=g)
        lens[0] = T
        mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
        ms_sm_eager = timed(lambda: enc(x, src_key_padding_mask=mask))
        ms_at_eager = timed(lambda: O.conformer_encoder_mhsa(x, sd, NL, H, act="swish", key_padding_mask=~mask))
        # both arms replayed as CUDA graphs (what a serving process does: no host launch work in the timed region)
        gf = S.GraphedForward(enc, x, mask)
        ms_sm = timed(gf.replay)
        try:
            nm = ~mask
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                O.conformer_encoder_mhsa(x, sd, NL, H, act="swish", key_padding_mask=nm)
            torch.cuda.current_stream(dev).wait_stream(side)
            ga = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                O.conformer_encoder_mhsa(x, sd, NL, H, act="swish", key_padding_mask=nm)
            ms_at = timed(ga.replay)
        except Exception:  # (capture is an optimisation of the host side only)
            ms_at = ms_at_eager
        del gf
        audio_s = B * T * 0.04
        rows.append({"T": T, "B": B, "summarymixing_ms": round(ms_sm, 3), "mhsa_torch_ms": round(ms_at, 3),
                     "summarymixing_eager_ms": round(ms_sm_eager, 3), "mhsa_torch_eager_ms": round(ms_at_eager, 3),
                     "summarymixing_frames_per_s": round(B * T / ms_sm * 1e3), "mhsa_torch_frames_per_s": round(B * T / ms_at * 1e3),
                     "summarymixing_rtf": ms_sm / 1e3 / audio_s, "mhsa_torch_rtf": ms_at / 1e3 / audio_s, "speedup": ms_at / ms_sm})
        print(f"T={T:5d}  SummaryMixing (libsmx) {ms_sm:8.3f} ms  RTF {rows[-1]['summarymixing_rtf']:.2e}   |   self-attention (torch, flash SDPA) "
              f"{ms_at:8.3f} ms  RTF {rows[-1]['mhsa_torch_rtf']:.2e}   |   x{ms_at / ms_sm:.2f}   (eager launches: {ms_sm_eager:.3f} / {ms_at_eager:.3f} ms)")
print(json.dumps({"config": "cfg5: utterance-length sweep, SummaryMixing (libsmx) vs self-attention Conformer (reference algorithm, torch CUDA ops)",
                  "batch": B, "layers": NL, "d_model": D, "io": "bf16", "rows": rows}))
