"""GPU parity of the bf16 tensor-core (tcgen05) arm against the CPU oracle at tile-aligned model dims.

The oracle is evaluated in fp32 on the SAME bf16-rounded input; the tensor-core arm additionally rounds
GEMM operands and inter-kernel activations to bf16 (fp32 accumulation and epilogues).  Tolerance:
max-abs <= 2.5e-2 * max(1, |y|max) and relative L2 <= 1.5e-2 — the floors measured for bf16 operands in
BASELINE.md section 5 (4e-3..1.2e-2 at |y|max 1.4-2.8), with headroom for the deeper compositions.
Every test also asserts that tcgen05 kernels actually ran (smx_tc_launch_count).
"""
import pytest
import torch
import torch.nn as nn

import summarymixing_b200 as S
from oracle import smx_oracle as O
from summarymixing_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _perturb(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in m.parameters():
            p.add_((0.02 if p.dim() >= 2 else 0.1) * torch.randn(p.shape, generator=g))


def _check(y, y_or, what, abs_tol=2.5e-2, rel_tol=1.5e-2):
    y = y.float().cpu()
    scale = max(1.0, float(y_or.abs().max()))
    err = float((y - y_or).abs().max())
    rel = float((y - y_or).norm() / y_or.norm())
    assert err <= abs_tol * scale and rel <= rel_tol, f"{what}: max-abs {err:.3e} (|y|max {scale:.2f}), rel-L2 {rel:.3e}"
    return err, rel


def _mask(B, T, lens):
    return torch.arange(T)[None] < torch.tensor(lens)[:, None]


@pytest.mark.parametrize("D,h,act", [(256, 4, "swish"), (256, 1, "gelu"), (128, 2, "swish"), (64, 1, "relu")])
def test_cell_tc_vs_oracle(D, h, act):
    acts = {"swish": S.Swish, "gelu": nn.GELU, "relu": nn.ReLU}
    torch.manual_seed(D + h)
    m = S.SummaryMixing(D, h, [D], D, [D], D, activation=acts[act]).eval()
    _perturb(m, 1)
    B, T = 3, 300
    x = torch.randn(B, T, D, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    mask = _mask(B, T, [300, 129, 5])
    y_or = O.summary_mixing(x.float(), dict(m.state_dict()), mode="SummaryMixing", act=act, src_padding_mask=mask)
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        y = m.to(DEV)(x.to(DEV), src_padding_mask=mask.to(DEV))
    torch.cuda.synchronize()
    expect = 1 if D in (128, 256) else 2  # one persistent kernel (K-SM v4); D=64: the two-pass v3 kernels
    assert L.lib().smx_tc_launch_count() - n0 == expect, "cell did not run on the fused tcgen05 kernel(s)"
    _check(y, y_or, f"cell D={D} h={h}")


@pytest.mark.parametrize("B,T,D,h,hid", [(32, 1000, 256, 4, 256), (5, 777, 256, 1, 128), (40, 130, 128, 2, 256), (2, 7, 64, 1, 64)])
def test_cell_fused_persistent_shapes(B, T, D, h, hid):
    """Persistent tile loop (more tiles than SMs), ragged last tiles, dense and block-diagonal weights, hidden != D."""
    torch.manual_seed(B + T)
    m = S.SummaryMixing(D, h, [hid], D, [hid], D, activation=S.Swish).eval()
    _perturb(m, 3)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None] < lens[:, None]
    y_or = O.summary_mixing(x.float(), dict(m.state_dict()), mode="SummaryMixing", act="swish", src_padding_mask=mask)
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        m = m.to(DEV)
        y = m(x.to(DEV), src_padding_mask=mask.to(DEV))
        y2 = m(x.to(DEV), src_padding_mask=mask.to(DEV))
    torch.cuda.synchronize()
    n_tiles = B * ((T + 127) // 128)
    one_kernel = D in (128, 256) and hid in (128, 256) and n_tiles <= 2 * torch.cuda.get_device_properties(0).multi_processor_count
    assert L.lib().smx_tc_launch_count() - n0 == (2 if one_kernel else 4)
    assert torch.equal(y, y2), "the fused cell must be run-to-run deterministic (fixed-order reductions)"
    _check(y, y_or, f"fused cell B={B} T={T} D={D} h={h}")


def test_cell_tc_no_mask_no_layernorm():
    torch.manual_seed(5)
    m = S.SummaryMixing(256, 4, [256], 256, [256], 256, activation=S.Swish, use_layernorm=False).eval()
    _perturb(m, 5)
    x = torch.randn(2, 128, 256, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16)
    y_or = O.summary_mixing(x.float(), dict(m.state_dict()), mode="SummaryMixing", act="swish", use_layernorm=False)
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        y = m.to(DEV)(x.to(DEV))
    assert L.lib().smx_tc_launch_count() > n0
    _check(y, y_or, "cell no-mask no-LN")


def test_conv_module_tc_vs_oracle():
    torch.manual_seed(7)
    m = S.ConvolutionModule(256, 31, True, S.Swish, 0.0, masked_false_or_true=False).eval()
    _perturb(m, 7)
    B, T = 3, 300
    x = torch.randn(B, T, 256, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    mask = _mask(B, T, [300, 150, 17])
    y_or = O.convolution_module(x.float(), dict(m.state_dict()), "", act="swish", mask=mask.unsqueeze(-1))
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        y = m.to(DEV)(x.to(DEV), mask.unsqueeze(-1).to(DEV))
    assert L.lib().smx_tc_launch_count() - n0 == 2, "conv module did not run on the tensor-core arm"
    _check(y, y_or, "conv module")


@pytest.mark.parametrize("D,F,h", [(256, 1024, 4), (128, 256, 2), (64, 128, 1), (192, 384, 3)])
def test_conformer_layer_tc_vs_oracle(D, F, h):
    torch.manual_seed(9)
    m = S.ConformerEncoderLayer(D, F, h, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval()
    _perturb(m, 9)
    B, T = 3, 300
    x = torch.randn(B, T, D, generator=torch.Generator().manual_seed(10)).to(torch.bfloat16)
    mask = _mask(B, T, [300, 211, 40])
    y_or = O.conformer_layer(x.float(), dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        y = m.to(DEV)(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    # 2 FFN + cell (1 launch for K-SM v4 at D=128/256, 2 passes otherwise) + 2 conv (D=192: unfused conv kernels, one more)
    expect = {256: 5, 128: 5, 64: 6, 192: 7}[D]
    assert L.lib().smx_tc_launch_count() - n0 in (expect, 6), "layer: expected 2 FFN + cell + 2 conv tcgen05 launches"
    _check(y, y_or, f"conformer layer D={D}", abs_tol=4e-2, rel_tol=2e-2)


def test_conformer_layer_persistent_bench_shape():
    """One layer at the bench shape (B=32, T=1000: 250/256 tiles > 148 SMs, so every persistent kernel loops)."""
    torch.manual_seed(21)
    D = 256
    m = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval()
    _perturb(m, 21)
    B, T = 32, 1000
    g = torch.Generator().manual_seed(22)
    x = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    lens = torch.randint(500, T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None] < lens[:, None]
    y_or = O.conformer_layer(x.float(), dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
    with torch.no_grad():
        m = m.to(DEV)
        y = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
        y2 = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    assert torch.equal(y, y2), "layer forward must be run-to-run deterministic"
    _check(y, y_or, "conformer layer at the bench shape", abs_tol=4e-2, rel_tol=2e-2)


def test_conformer_encoder_tc_vs_oracle_and_fp32_arm():
    """4 layers at D=256: the bf16 tensor-core arm vs the oracle, with the fp32 arm's error for reference."""
    torch.manual_seed(11)
    n = 4
    m = S.ConformerEncoder(n, 256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256],
                           local_proj_out_dim=256, summary_hid_dim=[256]).eval()
    _perturb(m, 11)
    B, T = 2, 257
    x = torch.randn(B, T, 256, generator=torch.Generator().manual_seed(12))
    mask = _mask(B, T, [257, 100])
    y_or = O.conformer_encoder(x, dict(m.state_dict()), n, act="swish", src_key_padding_mask=mask)
    m = m.to(DEV)
    with torch.no_grad():
        y32 = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
        n0 = L.lib().smx_tc_launch_count()
        y16 = m(x.to(torch.bfloat16).to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    assert L.lib().smx_tc_launch_count() - n0 == 5 * n  # per layer: 2 FFN + 1 cell + 2 conv-module kernels
    assert float((y32.cpu() - y_or).abs().max()) < 5e-4
    _check(y16, y_or, "conformer encoder (4 layers) bf16 tensor-core arm", abs_tol=6e-2, rel_tol=3e-2)


@pytest.mark.parametrize("version", [2, 3, 4])
def test_ffn_kernel_generations_agree(version):
    """K-FFN v2 (hidden chunk in shared memory), v3 (hidden chunk in tensor memory, A operand of GEMM2 read from TMEM)
    and v4 (v3 on CTA pairs, cta_group::2) compute the same layer: each against the oracle at the bench width, with
    more tiles than SMs and a ragged last tile."""
    torch.manual_seed(31)
    D = 256
    m = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval()
    _perturb(m, 31)
    B, T = 21, 933
    g = torch.Generator().manual_seed(32)
    x = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    lens = torch.randint(300, T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None] < lens[:, None]
    y_or = O.conformer_layer(x.float(), dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
    try:
        assert L.lib().smx_debug_set_ffn_version(version) == 0
        with torch.no_grad():
            y = m.to(DEV)(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
        torch.cuda.synchronize()
    finally:
        L.lib().smx_debug_set_ffn_version(4)  # the default
    _check(y, y_or, f"conformer layer, FFN generation {version}", abs_tol=4e-2, rel_tol=2e-2)


def test_cfg3_conformer_large_dims():
    """BASELINE.json configs[2]: the conformer_summarymixing.yaml dims (D=512, h=8, d_ffn=2048, hid/out 512) -- wider than
    the fused tcgen05 kernels take (D <= 256), so the layer runs on the unfused tensor-core kernels / the fp32-math arm.
    fp32 I/O within 5e-4 of the oracle, bf16 I/O within the bf16 tolerance."""
    torch.manual_seed(41)
    D = 512
    m = S.ConformerEncoderLayer(D, 2048, 8, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval()
    _perturb(m, 41)
    B, T = 2, 170
    x = torch.randn(B, T, D, generator=torch.Generator().manual_seed(42))
    mask = _mask(B, T, [170, 61])
    y_or = O.conformer_layer(x, dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
    m = m.to(DEV)
    with torch.no_grad():
        y32 = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
        xb = x.to(torch.bfloat16)
        y16 = m(xb.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    assert float((y32.cpu() - y_or).abs().max()) < 5e-4
    y_or16 = O.conformer_layer(xb.float(), dict(m.cpu().state_dict()), "", act="swish", src_key_padding_mask=mask)
    _check(y16, y_or16, "conformer_large layer (D=512) bf16 I/O", abs_tol=4e-2, rel_tol=2e-2)


def test_cfg4_branchformer_lite_dims():
    """BASELINE.json configs[3]: Branchformer SummaryMixing-lite, D=512, csgu 3072, ragged padded batch with mask;
    fp32 within 5e-4 of the oracle."""
    torch.manual_seed(43)
    D = 512
    m = S.BranchformerEncoderLayer(D, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[D], local_proj_out_dim=D,
                                   summary_hid_dim=[D], summary_out_dim=D, mode="SummaryMixing-lite").eval()
    _perturb(m, 43)
    B, T = 3, 230
    x = torch.randn(B, T, D, generator=torch.Generator().manual_seed(44))
    mask = _mask(B, T, [230, 200, 37])
    y_or = O.branchformer_layer(x, dict(m.state_dict()), "", mode="SummaryMixing-lite", src_key_padding_mask=mask)
    with torch.no_grad():
        y = m.to(DEV)(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    torch.cuda.synchronize()
    assert float((y.cpu() - y_or).abs().max()) < 5e-4


@pytest.mark.parametrize("version", [1, 3, 4])
def test_cell_kernel_generations_agree(version):
    """K-SM first generation (operands staged through shared memory) and v3 (H / L operands resident in tensor memory,
    schedule-ordered weight ring) compute the same cell and GLU pass: each against the oracle on a ragged batch with
    more tiles than SMs; run-to-run bit-identical."""
    torch.manual_seed(51)
    D = 256
    m = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval()
    _perturb(m, 51)
    B, T = 23, 901
    g = torch.Generator().manual_seed(52)
    x = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    lens = torch.randint(200, T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None] < lens[:, None]
    y_or = O.conformer_layer(x.float(), dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
    try:
        assert L.lib().smx_debug_set_cell_version(version) == 0
        with torch.no_grad():
            m = m.to(DEV)
            y = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
            y2 = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
        torch.cuda.synchronize()
    finally:
        L.lib().smx_debug_set_cell_version(4)
    assert torch.equal(y, y2)
    _check(y, y_or, f"conformer layer, cell generation {version}", abs_tol=4e-2, rel_tol=2e-2)


def test_programmatic_dependent_launch_is_transparent():
    """The fused kernels overlap their set-up (and pass B of the cell its first tile) with their predecessors through
    griddepcontrol; the results must be bit-identical with it switched off."""
    torch.manual_seed(61)
    n = 3
    m = S.ConformerEncoder(n, 256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256],
                           local_proj_out_dim=256, summary_hid_dim=[256]).eval().to(DEV)
    g = torch.Generator().manual_seed(62)
    B, T = 32, 1000
    x = torch.randn(B, T, 256, generator=g).to(torch.bfloat16).to(DEV)
    lens = torch.randint(500, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] < lens[:, None]).to(DEV)
    with torch.no_grad():
        outs = []
        for on in (1, 0, 1):
            try:
                assert L.lib().smx_debug_set_pdl(on) == 0
                for _ in range(3):  # back-to-back forwards: every kernel has a programmatic predecessor
                    y = m(x, src_key_padding_mask=mask)[0]
                torch.cuda.synchronize()
            finally:
                L.lib().smx_debug_set_pdl(1)
            outs.append(y.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_cuda_graph_replay_matches_eager():
    """The library is enqueue-only and allocation-free, and its programmatic-launch edges survive stream capture: a CUDA
    graph of the encoder forward (summarymixing_b200.GraphedForward) replays bit-identically to the eager call."""
    torch.manual_seed(71)
    m = S.ConformerEncoder(2, 256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256],
                           local_proj_out_dim=256, summary_hid_dim=[256]).eval().to(DEV)
    g = torch.Generator().manual_seed(72)
    B, T = 20, 700
    xs = [torch.randn(B, T, 256, generator=g).to(torch.bfloat16).to(DEV) for _ in range(2)]
    lens = torch.randint(100, T + 1, (B,), generator=g)
    masks = [(torch.arange(T)[None] < lens[:, None]).to(DEV), (torch.arange(T)[None] < lens.flip(0)[:, None]).to(DEV)]
    with torch.no_grad():
        eager = [m(x, src_key_padding_mask=k)[0].clone() for x, k in zip(xs, masks)]
        gf = S.GraphedForward(m, xs[0], masks[0])
        for i in (0, 1, 0):
            y = gf(xs[i], masks[i])
            torch.cuda.synchronize()
            assert torch.equal(y, eager[i])


def test_host_pipeline_serves_host_batches():
    """summarymixing_b200.HostPipeline is the package's host-buffer entry point (what bench.py's `e2e` number is measured
    through): pageable and pinned HOST batches in, pinned host results out, copies on their own streams around graph replays.
    Every result equals the module's forward on the same batch, bit for bit, in order; device tensors and wrong shapes are
    refused loudly, and the modules themselves keep refusing host tensors (no CPU path)."""
    torch.manual_seed(73)
    D, B, T = 256, 6, 400
    m = S.ConformerEncoder(2, D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                           local_proj_out_dim=D, summary_hid_dim=[D]).eval().to(DEV)
    g = torch.Generator().manual_seed(74)
    batches = []
    for i in range(5):
        x = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
        lens = torch.randint(50, T + 1, (B,), generator=g)
        mask = torch.arange(T)[None] < lens[:, None]
        if i % 2:  # every other batch already pinned: taken as is, the others are staged through the pipeline's own buffers
            x, mask = x.pin_memory(), mask.pin_memory()
        batches.append((x, mask))
    with torch.no_grad():
        want = [m(x.to(DEV), src_key_padding_mask=k.to(DEV))[0].cpu() for x, k in batches]
    pipe = S.HostPipeline(m, B, T, D, device=DEV)
    assert pipe.h2d_bytes_per_step == B * T * D * 2 + B * T and pipe.d2h_bytes_per_step == B * T * D * 2
    t0 = L.lib().smx_tc_launch_count()
    got = [y.clone() for y in pipe.run(batches)]
    assert len(got) == len(want)
    for y, w in zip(got, want):
        assert not y.is_cuda and torch.equal(y, w)
    if pipe.graphs is None:  # eager launches: the kernels of 5 forwards were enqueued by this call
        assert L.lib().smx_tc_launch_count() - t0 == 5 * 2 * 5
    with pytest.raises(RuntimeError, match="HOST tensors"):
        list(pipe.run([(batches[0][0].to(DEV), batches[0][1].to(DEV))]))
    with pytest.raises(RuntimeError, match="was built for"):
        list(pipe.run([(batches[0][0][:, :100], batches[0][1][:, :100])]))
    with pytest.raises(Exception):
        m(batches[0][0], src_key_padding_mask=batches[0][1])


@pytest.mark.parametrize("B,T,launches", [(1, 16, 5), (2, 127, 5), (1, 129, 5), (5, 257, 5), (37, 300, 5), (148, 256, 5), (149, 256, 6), (40, 1000, 6)])
def test_layer_edge_shapes_and_resident_kernel_limits(B, T, launches):
    """Edge shapes of the D = 256 layer on the tensor-core arm against the fp32-math arm of the same library: fewer rows than one
    tile (TMA boxes larger than the tensor), ragged last tiles, exactly two tiles per CTA (148 x 256 frames = 296 tiles) and one
    tile more -- where the one-kernel cell and K-GLU v4 (<= two resident tiles per CTA) hand over to their multi-pass fallbacks
    (the cell then takes two tensor-core launches)."""
    torch.manual_seed(5)
    D = 256
    layer = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                                    summary_hid_dim=[D]).eval().to(DEV)
    g = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, T, D, generator=g)
    lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None] < lens[:, None]).to(DEV)
    with torch.no_grad():
        y32 = layer(x.to(DEV), src_key_padding_mask=mask)[0]
        t0 = L.lib().smx_tc_launch_count()
        y16 = layer(x.to(torch.bfloat16).to(DEV), src_key_padding_mask=mask)[0].float()
        n = L.lib().smx_tc_launch_count() - t0
    torch.cuda.synchronize()
    valid = mask.unsqueeze(-1)
    rel = float(((y16 - y32) * valid).norm() / (y32 * valid).norm())
    assert n == launches, n
    assert torch.isfinite(y16).all() and rel < 1e-2, rel


def test_folded_norm1_agrees_with_the_in_place_layernorm():
    """smx_cell_pack_prenorm folds the layer's norm1 into the one-kernel cell's image (gamma into the first blocks' weights,
    statistics-only pass, row scalars in the first epilogue).  The layer records the folded parameters in its cell struct, the
    folded and the unfolded kernel (SMX_C4_NOFOLD=1: in-place LayerNorm of the tile) agree to bf16 rounding, and both stay
    within the module-level bound of the oracle."""
    import os

    torch.manual_seed(91)
    D, B, T = 256, 9, 700
    m = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                                summary_hid_dim=[D]).eval()
    _perturb(m, 91)
    g = torch.Generator().manual_seed(92)
    x = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    lens = torch.randint(200, T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None] < lens[:, None]
    y_or = O.conformer_layer(x.float(), dict(m.state_dict()), "", act="swish", src_key_padding_mask=mask)
    m = m.to(DEV)
    with torch.no_grad():
        y_fold = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0].float().cpu()
        lw = m._wv.struct
        assert lw.cell.prenorm_w == lw.norm1_w and lw.cell.prenorm_b == lw.norm1_b and lw.cell.prenorm_w
        os.environ["SMX_C4_NOFOLD"] = "1"
        try:
            y_plain = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0].float().cpu()
        finally:
            del os.environ["SMX_C4_NOFOLD"]
    rel = float((y_fold - y_plain).norm() / y_plain.norm())
    assert 0.0 < rel < 5e-3, rel  # different kernels (not bit-equal), same function
    _check(y_fold, y_or, "layer, norm1 folded", abs_tol=4e-2, rel_tol=2e-2)
    _check(y_plain, y_or, "layer, norm1 in place", abs_tol=4e-2, rel_tol=2e-2)


def test_16_byte_aligned_buffers_take_the_fallback_kernels():
    """The newest kernels move rows with 256-bit accesses and need 32-byte aligned activations; buffers that are only
    16-byte aligned (the C ABI's stated minimum) must still give the same answer through the fallbacks: first-generation
    cell / FFN kernels for an unaligned input, two 128-bit stores in the conv kernel for an unaligned output."""
    import ctypes as C

    from summarymixing_b200 import _host as H

    torch.manual_seed(81)
    D, B, T = 256, 5, 333
    m = S.ConformerEncoderLayer(D, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                                local_proj_out_dim=D, summary_hid_dim=[D]).eval().to(DEV)
    g = torch.Generator().manual_seed(82)
    x = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    lens = torch.randint(50, T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None] < lens[:, None])
    with torch.no_grad():
        y_ref = m(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
        # (1) the same input at an address that is 16- but not 32-byte aligned
        flat = torch.empty(B * T * D + 8, dtype=torch.bfloat16, device=DEV)
        xu = flat[8:].view(B, T, D)
        xu.copy_(x)
        assert xu.data_ptr() % 32 == 16 and xu.is_contiguous()
        y_u = m(xu, src_key_padding_mask=mask.to(DEV))[0]
    torch.cuda.synchronize()
    d = float((y_u.float() - y_ref.float()).abs().max())
    assert d <= 2.5e-2 * max(1.0, float(y_ref.float().abs().max())), d  # different kernel generation on the first FFN: bf16 tolerance

    # (2) the conv module straight through the C ABI with an output pointer that is only 16-byte aligned
    lib = L.lib()
    with torch.no_grad():
        m(x.to(DEV), src_key_padding_mask=mask.to(DEV))  # make sure the weight structs are filled
    lw = m._wv.struct
    m8 = mask.to(torch.uint8).to(DEV).contiguous()
    xd = x.to(DEV).contiguous()
    nb = max(lib.smx_conv_module_workspace_bytes(C.byref(lw.conv), L.BF16, B, T), 1 << 20)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    st = H.stream_ptr(torch.device(DEV))
    ya = torch.empty(B * T * D + 8, dtype=torch.bfloat16, device=DEV)
    yb = torch.empty(B * T * D + 8, dtype=torch.bfloat16, device=DEV)
    assert ya.data_ptr() % 32 == 0
    L.check(lib.smx_conv_module_fwd(C.byref(lw.conv), lw.act, L.BF16, B, T, 0, xd.data_ptr(), m8.data_ptr(), xd.data_ptr(),
                                    ya.data_ptr(), ws.data_ptr(), ws.numel(), st))
    L.check(lib.smx_conv_module_fwd(C.byref(lw.conv), lw.act, L.BF16, B, T, 0, xd.data_ptr(), m8.data_ptr(), xd.data_ptr(),
                                    yb.data_ptr() + 16, ws.data_ptr(), ws.numel(), st))
    torch.cuda.synchronize()
    assert torch.equal(ya[: B * T * D], yb[8: 8 + B * T * D]), "aligned and unaligned output stores must give identical bytes"


def test_tensor_core_arm_is_bit_exactly_batch_invariant():
    """Size-independent property at the bench shape (B=32, T=1000, D=256, bf16 tensor-core arm): the path is per utterance,
    so permuting the batch, or running a slice of it alone (different tile -> CTA assignment, different tile positions in
    the flat row space of the FFN kernels), must reproduce the same rows BIT FOR BIT -- the partition that multi-GPU
    sharding relies on."""
    torch.manual_seed(91)
    m = S.ConformerEncoder(2, 256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256],
                           local_proj_out_dim=256, summary_hid_dim=[256]).eval().to(DEV)
    g = torch.Generator().manual_seed(92)
    B, T = 32, 1000
    x = torch.randn(B, T, 256, generator=g).to(torch.bfloat16).to(DEV)
    lens = torch.randint(500, T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None] < lens[:, None]).to(DEV)
    with torch.no_grad():
        y = m(x, src_key_padding_mask=mask)[0]
        perm = torch.randperm(B, generator=g).to(DEV)
        yp = m(x[perm].contiguous(), src_key_padding_mask=mask[perm].contiguous())[0]
        ys = m(x[5:18].contiguous(), src_key_padding_mask=mask[5:18].contiguous())[0]
        y1 = m(x[31:32].contiguous(), src_key_padding_mask=mask[31:32].contiguous())[0]
    torch.cuda.synchronize()
    assert torch.equal(yp, y[perm])
    assert torch.equal(ys, y[5:18])
    assert torch.equal(y1, y[31:32])


def test_invalidate_weights_after_a_data_write():
    """A write through .data does not bump the version counter: the cached packed images stay; invalidate_weights makes the next call
    re-read the parameters."""
    import summarymixing_b200 as S

    torch.manual_seed(3)
    m = S.SummaryMixing(256, 4, [256], 256, [256], 256, activation=S.Swish).to("cuda:0").eval()
    x = torch.randn(2, 200, 256, device="cuda:0").bfloat16()
    with torch.no_grad():
        y0 = m(x).clone()
        m.summary_local_merging["linear"].w.weight.data.mul_(1.5)   # (a GEMM weight: the tensor-core arm reads its packed bf16 image)
        y_cached = m(x).clone()
        assert torch.equal(y_cached, y0)          # (documented behaviour: the write was invisible to the cache)
        assert S.invalidate_weights(m) >= 1
        y1 = m(x)
    assert float((y1.float() - y0.float()).abs().max()) > 1e-2
