"""GPU parity: the CUDA path (through the module surface -> C ABI -> libsmx kernels) against
(1) the committed golden vectors produced by the unmodified reference, and (2) the CPU oracle on seeded
inputs, plus size-independent properties at the BASELINE shape.

Tolerances (max-abs):
  fp32 I/O  : 1e-4 cell-level, 5e-4 encoder-level (fp32 arithmetic, different summation order; the
              reference's own fp32-vs-fp64 gap on these cases is ~1e-6).
  bf16 I/O  : the error of rounding input/output to bf16 dominates: bounded by 3e-2 on |y|<=6 outputs,
              and always compared against the error the oracle itself makes when fed bf16-rounded input.
"""
import pytest
import torch

from oracle import smx_oracle as O
from tests import _golden as G
from tests._build import module_from_fixture, run_module
from tests.test_oracle_golden import run_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tol(kind, dtype):
    enc = kind in ("conformer_layer", "conformer_encoder", "branchformer_encoder")
    if dtype == torch.float32:
        return 5e-4 if enc else 1e-4
    return 6e-2 if enc else 3e-2


@pytest.mark.parametrize("name", G.names())
def test_fp32_matches_reference_golden(name):
    fx = G.Fixture(name)
    m = module_from_fixture(fx).to(DEV)
    with torch.no_grad():
        y = run_module(m, fx, fx.x.to(DEV), DEV)
    torch.cuda.synchronize()
    assert tuple(y.shape) == tuple(fx.y.shape)
    err = float((y.float().cpu() - fx.y).abs().max())
    assert err < _tol(fx.cfg["kind"], torch.float32), f"{name}: max-abs {err:.3e}"


@pytest.mark.parametrize("name", G.names())
def test_bf16_io_matches_reference_golden(name):
    fx = G.Fixture(name)
    m = module_from_fixture(fx).to(DEV)
    xb = fx.x.to(torch.bfloat16)
    with torch.no_grad():
        y = run_module(m, fx, xb.to(DEV), DEV)
    torch.cuda.synchronize()
    assert y.dtype == torch.bfloat16
    # what the oracle gives for the same bf16-rounded input, in fp32
    fx.x = xb.float()
    y_or = run_oracle(fx, torch.float32)
    err = float((y.float().cpu() - y_or).abs().max())
    scale = max(1.0, float(y_or.abs().max()))
    assert err < 1.2e-2 * scale, f"{name}: bf16 max-abs {err:.3e} (|y|max {scale:.2f})"


def test_loads_native_library():
    import summarymixing_b200._lib as L

    before = L.lib().smx_launch_count()
    fx = G.Fixture("cell_sm_h4_swish")
    m = module_from_fixture(fx).to(DEV)
    with torch.no_grad():
        run_module(m, fx, fx.x.to(DEV), DEV)
    assert L.lib().smx_launch_count() > before


# ---- edge cases the reference's behaviour defines ----------------------------------------------
def test_default_mask_equals_all_ones_mask():
    fx = G.Fixture("cell_sm_h4_swish")
    m = module_from_fixture(fx).to(DEV)
    x = fx.x.to(DEV)
    with torch.no_grad():
        y0 = m(x)
        y1 = m(x, src_padding_mask=torch.ones(x.shape[:2], device=DEV))
        y2 = m(x, src_padding_mask=torch.ones(x.shape[:2], device=DEV, dtype=torch.bool))
    assert torch.equal(y0, y1) and torch.equal(y0, y2)


def test_padded_frames_are_constant_nonzero():
    fx = G.Fixture("cell_sm_h4_swish")
    m = module_from_fixture(fx).to(DEV)
    with torch.no_grad():
        y = m(fx.x.to(DEV), src_padding_mask=fx.mask.to(DEV))
    pad = y[3, 7:]
    assert float(pad.abs().max()) > 0
    assert float((pad - pad[:1]).abs().max()) < 1e-6


def test_lite_returns_expand_view():
    fx = G.Fixture("cell_sm_lite_h4_gelu")
    m = module_from_fixture(fx).to(DEV)
    with torch.no_grad():
        y = m(fx.x.to(DEV), src_padding_mask=fx.mask.to(DEV))
    assert y.stride(1) == 0 and tuple(y.shape) == tuple(fx.y.shape)


def test_single_frame_and_single_utterance():
    fx = G.Fixture("cell_sm_h4_swish")
    m = module_from_fixture(fx).to(DEV)
    x = fx.x[:1, :1].contiguous()
    with torch.no_grad():
        y = m(x.to(DEV))
    y_or = O.summary_mixing(x, fx.sd, mode="SummaryMixing", act="swish")
    assert float((y.cpu() - y_or).abs().max()) < 1e-4


def test_all_padded_utterance_gives_nan_like_reference():
    # sum(mask)=0 -> 0/0 in the reference (summary_mixing.py:229-231): NaN for that utterance only
    fx = G.Fixture("cell_sm_h4_swish")
    m = module_from_fixture(fx).to(DEV)
    mask = fx.mask.clone()
    mask[2] = False
    with torch.no_grad():
        y = m(fx.x.to(DEV), src_padding_mask=mask.to(DEV)).cpu()
    y_or = O.summary_mixing(fx.x, fx.sd, mode="SummaryMixing", act="swish", src_padding_mask=mask)
    assert torch.isnan(y_or[2]).all() and torch.isnan(y[2]).all()
    assert float((y[[0, 1, 3]] - y_or[[0, 1, 3]]).abs().max()) < 1e-4


def test_ragged_odd_sizes_against_oracle():
    """Sizes that are not multiples of any tile: D=72 (h=3), T=131, B=5, ragged lengths."""
    import summarymixing_b200 as S
    import torch.nn as nn

    torch.manual_seed(3)
    m = S.SummaryMixing(72, 3, [48], 60, [96], 36, activation=nn.GELU).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(5, 131, 72)
    lens = torch.tensor([131, 1, 77, 130, 64])
    mask = torch.arange(131)[None] < lens[:, None]
    y_or = O.summary_mixing(x, dict(m.state_dict()), mode="SummaryMixing", act="gelu", src_padding_mask=mask)
    with torch.no_grad():
        y = m.to(DEV)(x.to(DEV), src_padding_mask=mask.to(DEV))
    assert float((y.cpu() - y_or).abs().max()) < 1e-4


def test_mask_builders_match_reference_golden():
    import numpy as np
    from summarymixing_b200.lobes.models.transformer.TransformerASR import (
        make_transformer_src_mask,
        make_transformer_src_tgt_masks,
    )
    from tests._build import _DC

    z = np.load(G.GOLDEN_DIR + "/masks.npz")
    src = torch.zeros(4, 37, 8, device=DEV)
    pad, _, _, _ = make_transformer_src_tgt_masks(src, None, torch.from_numpy(z["wav_len"]), masked_false_or_true=False)
    assert np.array_equal(pad.cpu().numpy(), z["padding_mask"])
    for key in z.files:
        if key.startswith("chunk_"):
            _, cs, lc = key.split("_")
            m = make_transformer_src_mask(src, False, False, _DC(int(cs), None if lc == "None" else int(lc)))
            assert np.array_equal(m.cpu().numpy(), z[key]), key


# ---- BASELINE shape: size-independent properties -----------------------------------------------
def _baseline_encoder(n_layers=2):
    import summarymixing_b200 as S

    torch.manual_seed(0)
    enc = S.ConformerEncoder(n_layers, 256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256],
                             local_proj_out_dim=256, summary_hid_dim=[256], mode="SummaryMixing").eval()
    return enc


def test_baseline_shape_utterance_independence_and_batch_invariance():
    """(B=32,T=1000,D=256): the path is per-utterance — running utterances alone, or permuting the
    batch, must give the same rows (linearity of the batch partition used for multi-GPU sharding)."""
    enc = _baseline_encoder(2).to(DEV)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(32, 1000, 256, generator=g).to(DEV)
    lens = torch.randint(500, 1001, (32,), generator=g)
    lens[0] = 1000
    mask = (torch.arange(1000)[None] < lens[:, None]).to(DEV)
    with torch.no_grad():
        y = enc(x, src_key_padding_mask=mask)[0]
        perm = torch.randperm(32, generator=g).to(DEV)
        yp = enc(x[perm].contiguous(), src_key_padding_mask=mask[perm].contiguous())[0]
        y5 = enc(x[5:6].contiguous(), src_key_padding_mask=mask[5:6].contiguous())[0]
    assert torch.isfinite(y).all()
    assert float((yp - y[perm]).abs().max()) < 1e-5
    assert float((y5 - y[5:6]).abs().max()) < 1e-5


def test_baseline_shape_one_layer_against_oracle_sample():
    """One D=256 layer at (B=4,T=1000): CUDA fp32 vs the CPU oracle (a few seconds on CPU)."""
    import summarymixing_b200 as S

    torch.manual_seed(1)
    layer = S.ConformerEncoderLayer(256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256],
                                    local_proj_out_dim=256, summary_hid_dim=[256]).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 1000, 256, generator=g)
    lens = torch.tensor([1000, 873, 500, 641])
    mask = torch.arange(1000)[None] < lens[:, None]
    y_or = O.conformer_layer(x, dict(layer.state_dict()), "", act="swish", src_key_padding_mask=mask)
    with torch.no_grad():
        y = layer.to(DEV)(x.to(DEV), src_key_padding_mask=mask.to(DEV))[0]
    assert float((y.cpu() - y_or).abs().max()) < 5e-4


def test_sum_mask_chunk_prefix_path_and_general_fallback():
    """Per-frame summaries with a (T,T) sum mask (summary_mixing.py:235-246).  Dynamic-chunk masks (rows = runs of ones,
    TransformerASR.py:85-110) take the chunk-prefix-sum path (O(T D)); an arbitrary weighted matrix must still give the
    reference's (T,T) @ (T,D) result through the fallback product.  fp32 arm vs the oracle at T = 1000."""
    import summarymixing_b200 as S
    from oracle import smx_oracle as O
    from oracle.seeded import fill_module, seeded_input

    B, T, D = 3, 1000, 64
    m = S.SummaryMixing(D, 4, [D], D, [D], D, activation=S.Swish, mode="SummaryMixing").eval()
    fill_module(m, 91)
    x = seeded_input(92, B, T, D)
    lens = torch.tensor([1000, 731, 40])
    mask = torch.arange(T)[None] < lens[:, None]
    chunk = O.chunk_mask(T, 16, 4).float()                       # 16-frame chunks, 4 chunks of left context
    weighted = torch.rand(T, T, generator=torch.Generator().manual_seed(93)) + 0.1
    holes = chunk.clone()
    holes[5, 2] = 0.0                                            # ones with a hole: not an interval -> fallback
    md = m.to("cuda:0")
    for name, sm in (("chunk", chunk), ("weighted", weighted), ("holes", holes)):
        y_or = O.summary_mixing(x, dict(m.cpu().state_dict()), mode="SummaryMixing", act="swish", src_padding_mask=mask, sum_mask=sm)
        with torch.no_grad():
            y = md.to("cuda:0")(x.to("cuda:0"), sum_mask=sm.to("cuda:0"), src_padding_mask=mask.to("cuda:0")).cpu()
        err = float((y - y_or).abs().max())
        assert err <= 5e-4 * max(1.0, float(y_or.abs().max())), f"sum_mask {name}: {err:.3e}"
