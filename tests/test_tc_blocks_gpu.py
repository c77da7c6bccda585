"""GPU: the tcgen05 building blocks (operand layouts, UMMA descriptors, bulk copies, TMEM read-back)
checked through the diagnostic C-ABI entry smx_debug_tc_gemm against a plain fp32 matmul of the same
bf16-rounded operands (products exact in fp32; only the summation order differs: tol 1e-3 at K<=256)."""
import pytest
import torch

from summarymixing_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(layout, M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(DEV)
    w = torch.randn(N, K, generator=g).to(DEV)
    c = torch.full((M, N), float("nan"), device=DEV)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    L.check(L.lib().smx_debug_tc_gemm(layout, M, N, K, a.data_ptr(), w.data_ptr(), c.data_ptr(), ws.data_ptr(),
                                      ws.numel(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = a.float() @ w.to(torch.bfloat16).float().T
    return float((c - ref).abs().max()), float(ref.abs().max())


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("shape", [(128, 64, 64), (128, 256, 256), (300, 128, 192), (77, 48, 64), (256, 200, 128),
                                   (128, 16, 16)])
def test_umma_gemm_matches_matmul(layout, shape):
    M, N, K = shape
    err, scale = _run(layout, M, N, K)
    assert err < 1e-3 * max(1.0, scale), f"layout {layout} {shape}: max-abs {err:.3e} (|ref|max {scale:.1f})"
