"""Parity of the MEASURED configurations at their real sizes (BASELINE.json configs[1], [2], [3]).

The kernel runs with bf16 I/O on the tensor-core arm where it exists; it is compared with the CPU oracle in fp32 on the
same bf16-representable input, and -- in the same test -- with the reference ALGORITHM'S OWN bf16-vs-fp32 error (the oracle
evaluated with bfloat16 weights and activations, i.e. what the reference module does after .bfloat16(); the oracle is
pinned to the unmodified reference at these model dims by tests/test_tile_golden.py).  Asserted: kernel error <= that error
(SURVEY.md 8d), max-abs and relative L2; all numbers are printed.
"""
import time

import pytest
import torch

import summarymixing_b200 as S
from oracle import smx_oracle as O
from oracle.seeded import fill_module, seeded_input
from summarymixing_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _lens_mask(B, T, lo, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(lo, T + 1, (B,), generator=g)
    lens[0] = T
    return torch.arange(T)[None] < lens[:, None]


def _compare(tag, y, y32, y16):
    y = y.float().cpu()
    err, rel = float((y - y32).abs().max()), float((y - y32).norm() / y32.norm())
    rerr, rrel = float((y16 - y32).abs().max()), float((y16 - y32).norm() / y32.norm())
    print(f"\n[{tag}] |y|max {float(y32.abs().max()):.2f}  kernel: max-abs {err:.3e} rel-L2 {rel:.3e}  |  "
          f"reference algorithm in bf16: max-abs {rerr:.3e} rel-L2 {rrel:.3e}")
    assert rel <= rrel, f"{tag}: rel-L2 {rel:.3e} > reference-bf16 {rrel:.3e}"
    assert err <= rerr, f"{tag}: max-abs {err:.3e} > reference-bf16 {rerr:.3e}"
    return err, rel, rerr, rrel


def _oracle_pair(fn, sd, x):
    t0 = time.time()
    y32 = fn(x, sd)
    sd16 = {k: v.to(torch.bfloat16) if v.dtype.is_floating_point else v for k, v in sd.items()}
    y16 = fn(x.to(torch.bfloat16), sd16).float()
    print(f"(oracle fp32 + bf16: {time.time() - t0:.1f} s on {torch.get_num_threads()} threads)")
    return y32, y16


def test_cfg2_encoder_12_layers_bench_shape():
    """BASELINE configs[1]: the 12-layer D=256 encoder at B=32, T=1000 -- the exact shape bench.py times."""
    D, F, h, NL, B, T = 256, 1024, 4, 12, 32, 1000
    enc = S.ConformerEncoder(NL, D, F, h, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                             summary_hid_dim=[D], mode="SummaryMixing").eval()
    fill_module(enc, 51)
    x = seeded_input(52, B, T, D)
    mask = _lens_mask(B, T, 500, 53)
    sd = dict(enc.state_dict())
    y32, y16 = _oracle_pair(lambda xx, s: O.conformer_encoder(xx, s, NL, act="swish", src_key_padding_mask=mask), sd, x)
    enc = enc.to(DEV)
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        y = enc(x.to(torch.bfloat16).to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    torch.cuda.synchronize()
    tc = L.lib().smx_tc_launch_count() - n0
    assert tc >= NL * 5, f"the encoder did not run on the fused tcgen05 kernels ({tc} launches)"
    _compare(f"cfg2 12L B={B} T={T} (bf16 tcgen05 arm, {tc} tcgen05 launches)", y, y32, y16)
    # the fp32-I/O arm on the same input meets north_star's 1e-3 ON THE TENSOR CORES: its linears run as split-bf16 GEMMs
    # (hi*hi + hi*lo + lo*hi in fp32 accumulators, smx_tc_gemm.cu); with the switch off every product is an fp32 FMA
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        yf = enc(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0].cpu()
    tc32 = L.lib().smx_tc_launch_count() - n0
    e32 = float((yf - y32).abs().max())
    try:
        L.lib().smx_debug_set_f32_tc(0)
        with torch.no_grad():
            ye = enc(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0].cpu()
    finally:
        L.lib().smx_debug_set_f32_tc(1)
    ee = float((ye - y32).abs().max())
    print(f"[cfg2 12L fp32-I/O arm] split-bf16 tensor-core linears ({tc32} tcgen05 launches): max-abs {e32:.3e}; CUDA-core fp32 GEMMs: {ee:.3e}")
    assert tc32 >= NL * 8, "the fp32 arm's linears did not run on the tensor cores"
    assert e32 <= 1e-3 * max(1.0, float(y32.abs().max()))
    assert ee <= 1e-4 * max(1.0, float(y32.abs().max()))


def test_cfg3_conformer_large_encoder_real_dims():
    """BASELINE configs[2] model (conformer_summarymixing.yaml:113-125): 12 layers, D=512, h=8, d_ffn=2048, B=8, T=1000."""
    D, F, h, NL, B, T = 512, 2048, 8, 12, 8, 1000
    enc = S.ConformerEncoder(NL, D, F, h, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                             summary_hid_dim=[D], mode="SummaryMixing").eval()
    fill_module(enc, 61)
    x = seeded_input(62, B, T, D)
    mask = _lens_mask(B, T, 400, 63)
    sd = dict(enc.state_dict())
    y32, y16 = _oracle_pair(lambda xx, s: O.conformer_encoder(xx, s, NL, act="swish", src_key_padding_mask=mask), sd, x)
    enc = enc.to(DEV)
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        y = enc(x.to(torch.bfloat16).to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    torch.cuda.synchronize()
    tc = L.lib().smx_tc_launch_count() - n0
    assert tc > 0, "cfg3: no tcgen05 kernel ran"
    _compare(f"cfg3 12L D=512 B={B} T={T} ({tc} tcgen05 launches)", y, y32, y16)


def test_cfg4_branchformer_lite_real_dims():
    """BASELINE configs[3] model (branchformer_summarymixing.yaml:112-127, mode lite): 18 layers, D=512, csgu 3072,
    variable-length padded batch B=8, T=1200 (lengths 200..1200)."""
    D, NL, B, T = 512, 18, 8, 1200
    enc = S.BranchformerEncoder(NL, D, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[D], local_proj_out_dim=D,
                                summary_hid_dim=[D], summary_out_dim=D, mode="SummaryMixing-lite").eval()
    fill_module(enc, 71)
    x = seeded_input(72, B, T, D)
    mask = _lens_mask(B, T, 200, 73)
    sd = dict(enc.state_dict())
    y32, y16 = _oracle_pair(lambda xx, s: O.branchformer_encoder(xx, s, NL, act="gelu", gate_act="identity", mode="SummaryMixing-lite",
                                                                 src_key_padding_mask=mask), sd, x)
    enc = enc.to(DEV)
    with torch.no_grad():
        y = enc(x.to(torch.bfloat16).to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    torch.cuda.synchronize()
    _compare(f"cfg4 18L Branchformer-lite D=512 B={B} T={T}", y, y32, y16)
