"""Tile-aligned golden fixtures (D = 256 / 512, generated from the unmodified reference by oracle/gen_golden_tile.py).

CPU part (not gpu): pins the oracle -- and this repo's state_dict naming / seeded weight fill -- against the reference's
fp32 output at the sizes the fused tcgen05 kernels run at.
GPU part: the fp32-math arm against the reference output (<= 5e-4), and the bf16 tensor-core arm against it with the bound
SURVEY.md 8d sanctions: error <= the REFERENCE'S OWN bf16-vs-fp32 error on the same inputs (y16 in the fixture is the
unmodified reference module run in bfloat16).  Both numbers and the relative L2 error are printed side by side.
"""
import pytest
import torch

from tests import _golden as G
from tests import _models as M
from summarymixing_b200 import _lib as L

DEV = "cuda:0"
ALL = G.tile_names()
# fixtures whose bf16 path must run on tcgen05 kernels (fused path): everything at D = 256 in mode "SummaryMixing"
TC_REQUIRED = {"cell_d256_h4_swish", "cell_d256_h1_swish", "cell_d256_h4_gelu", "convmod_d256", "conformer_layer_d256",
               "conformer_enc_d256_3l", "cell_d256_h4_lite", "cell_d256_h4_fast", "conformer_layer_d512", "branchformer_enc_d512_lite_2l"}


def test_tile_fixture_inventory():
    assert len(ALL) >= 10, ALL


@pytest.mark.parametrize("name", ALL)
def test_oracle_matches_reference_at_tile_sizes(name):
    fx = G.TileFixture(name)
    m = M.build(fx.cfg)  # this repo's module only provides the state_dict (names + seeded weights); no CUDA involved
    y = M.run_oracle(fx.cfg, dict(m.state_dict()), fx.x, fx.mask)
    err = float((y - fx.y32).abs().max())
    assert err <= 2e-5 * max(1.0, float(fx.y32.abs().max())), f"{name}: oracle vs reference fp32 {err:.3e}"
    # the oracle in bfloat16 reproduces the reference-in-bfloat16 error level (same op sequence, same roundings)
    sd16 = {k: v.to(torch.bfloat16) if v.dtype.is_floating_point else v for k, v in m.state_dict().items()}
    y16 = M.run_oracle(fx.cfg, sd16, fx.x.to(torch.bfloat16), fx.mask).float()
    ref_abs, ref_rel = fx.ref_bf16_error()
    rel = float((y16 - fx.y32).norm() / fx.y32.norm())
    assert 0.5 * ref_rel <= rel <= 2.0 * ref_rel, f"{name}: oracle-bf16 rel-L2 {rel:.3e} vs reference-bf16 {ref_rel:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL)
def test_fp32_arm_matches_reference_at_tile_sizes(name):
    fx = G.TileFixture(name)
    m = M.build(fx.cfg).to(DEV)
    with torch.no_grad():
        y = M.run_module(m, fx.cfg, fx.x.to(DEV), fx.mask.to(DEV)).float().cpu()
    err = float((y - fx.y32).abs().max())
    assert err <= 5e-4 * max(1.0, float(fx.y32.abs().max())), f"{name}: fp32 arm vs reference {err:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ALL)
def test_bf16_arm_within_reference_bf16_error(name):
    fx = G.TileFixture(name)
    m = M.build(fx.cfg).to(DEV)
    n0 = L.lib().smx_tc_launch_count()
    with torch.no_grad():
        y = M.run_module(m, fx.cfg, fx.x.to(torch.bfloat16).to(DEV), fx.mask.to(DEV)).float().cpu()
    torch.cuda.synchronize()
    tc = L.lib().smx_tc_launch_count() - n0
    err, rel = float((y - fx.y32).abs().max()), float((y - fx.y32).norm() / fx.y32.norm())
    ref_abs, ref_rel = fx.ref_bf16_error()
    print(f"\n[{name}] |y|max {fx.cfg['y_absmax']:.2f}  kernel(bf16 I/O): max-abs {err:.3e} rel-L2 {rel:.3e}  |  "
          f"reference in bf16: max-abs {ref_abs:.3e} rel-L2 {ref_rel:.3e}  |  tcgen05 launches {tc}")
    if name in TC_REQUIRED:
        assert tc > 0, f"{name}: the bf16 path did not run on the tcgen05 arm"
    assert rel <= ref_rel, f"{name}: rel-L2 {rel:.3e} > reference-bf16 {ref_rel:.3e}"
    assert err <= ref_abs, f"{name}: max-abs {err:.3e} > reference-bf16 {ref_abs:.3e}"
