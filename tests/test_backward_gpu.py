"""GPU parity of smx_summary_mixing_bwd (through SummaryMixing.forward + autograd -> C ABI -> libsmx kernels) against
(1) gradients the unmodified reference produced (tests/golden/bwd, oracle/gen_golden_bwd.py) and (2) torch.autograd of
the CPU oracle on seeded inputs, plus size-independent properties at the BASELINE shape.

Tolerance (fp32 I/O): max-abs <= 1e-4 x max(1, |reference gradient|max) — fp32 arithmetic in a different summation order
(split-K over row slices); bf16 I/O: x, dy, dx rounded to bf16, compared against the oracle fed the same rounded inputs,
dx within 2 bf16 ulps of its scale, parameter gradients (fp32) within 1e-3 relative."""
import pytest
import torch
import torch.nn as nn

from oracle import smx_oracle as O
from tests import _golden as G
from tests._build import module_from_fixture, run_module
from tests.test_backward_golden import bwd_names, load_bwd

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, rel, what):
    err = float((a.detach().cpu().double() - b.double()).abs().max())
    assert err <= rel * max(1.0, float(b.abs().max())), f"{what}: max-abs {err:.3e} (reference max {float(b.abs().max()):.3e})"


@pytest.mark.parametrize("name", bwd_names())
def test_backward_matches_reference_gradients(name):
    fx = G.Fixture(name)
    dy, dx_ref, g_ref = load_bwd(name)
    m = module_from_fixture(fx).to(DEV)
    x = fx.x.to(DEV).requires_grad_(True)
    y = run_module(m, fx, x, DEV)
    assert y.requires_grad
    enc = fx.cfg["kind"] != "cell"
    _close(y, fx.y, 5e-4 if enc else 1e-4, "forward")
    y.backward(dy.to(DEV))
    tol = 3e-4 if enc else 1e-4  # deeper chains: fp32 rounding of more, longer sums
    _close(x.grad, dx_ref, tol, "dx")
    named = dict(m.named_parameters())
    assert set(g_ref) <= set(named)
    for k, v in g_ref.items():
        assert named[k].grad is not None, k
        _close(named[k].grad, v, tol, k)


def test_encoder_training_step_and_eval_path_agree():
    """Training mode (dropout=0): forward through the autograd chain equals the fused inference call, gradients reach
    every parameter, and an SGD step changes the output."""
    import summarymixing_b200 as S

    torch.manual_seed(21)
    enc = S.ConformerEncoder(2, 64, 128, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[64], local_proj_out_dim=64,
                             summary_hid_dim=[64], dropout=0.0).to(DEV)
    x = torch.randn(3, 90, 64, device=DEV)
    mask = (torch.arange(90)[None] < torch.tensor([90, 50, 17])[:, None]).to(DEV)
    with torch.no_grad():
        y_eval = enc.eval()(x, src_key_padding_mask=mask)[0]
    enc.train()
    y = enc(x, src_key_padding_mask=mask)[0]
    assert float((y.detach() - y_eval).abs().max()) < 1e-4
    loss = (y * mask[..., None]).pow(2).mean()
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc.parameters())
    opt = torch.optim.SGD(enc.parameters(), lr=0.05)
    opt.step()
    y2 = enc(x, src_key_padding_mask=mask)[0]
    loss2 = (y2 * mask[..., None]).pow(2).mean()
    assert float(loss2) < float(loss)


def _oracle_grads(m, x, mask, dy, act):
    sd = {k: v.detach().cpu().float().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xo = x.detach().cpu().float().clone().requires_grad_(True)
    y = O.summary_mixing(xo, sd, mode="SummaryMixing", act=act, use_layernorm=m.use_layernorm,
                         src_padding_mask=None if mask is None else mask.cpu())
    y.backward(dy.detach().cpu().float())
    return xo.grad, {k: v.grad for k, v in sd.items()}


def _perturbed(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
    return m


def test_backward_ragged_odd_sizes_against_oracle():
    """Nothing is a multiple of a tile: D=72 (h=3), hidden 48 / 96, D_l=60, D_s=36, T=131, ragged lengths; training mode
    (global_dropout=0) with only the parameters requiring grad."""
    import summarymixing_b200 as S

    torch.manual_seed(11)
    m = _perturbed(S.SummaryMixing(72, 3, [48], 60, [96], 36, activation=nn.GELU, global_dropout=0.0), 11).to(DEV).train()
    x = torch.randn(5, 131, 72, device=DEV)
    lens = torch.tensor([131, 1, 77, 130, 64])
    mask = (torch.arange(131)[None] < lens[:, None]).to(DEV)
    dy = torch.randn(5, 131, 36, device=DEV)
    y = m(x, src_padding_mask=mask)
    y.backward(dy)
    assert x.grad is None
    _, g = _oracle_grads(m, x, mask, dy, "gelu")
    for k, p in m.named_parameters():
        _close(p.grad, g[k], 1e-4, k)


def test_backward_bf16_io():
    import summarymixing_b200 as S

    torch.manual_seed(12)
    m = _perturbed(S.SummaryMixing(64, 4, [64], 64, [64], 64, activation=S.Swish), 12).to(DEV).eval()
    x = torch.randn(3, 200, 64, device=DEV).bfloat16().requires_grad_(True)
    mask = (torch.arange(200)[None] < torch.tensor([200, 150, 9])[:, None]).to(DEV)
    dy = torch.randn(3, 200, 64, device=DEV).bfloat16()
    m(x, src_padding_mask=mask).backward(dy)
    assert x.grad.dtype == torch.bfloat16
    dx, g = _oracle_grads(m, x, mask, dy, "swish")
    _close(x.grad.float(), dx, 2 ** -7, "dx (bf16)")
    for k, p in m.named_parameters():
        assert p.grad.dtype == torch.float32
        _close(p.grad, g[k], 1e-3, k)


def test_backward_properties_at_baseline_shape():
    """B=32, T=1000, D=256, h=4 (BASELINE configs[1] cell): the backward is linear in dy, padded frames receive no
    gradient, and an utterance's dx does not depend on the other utterances in the batch."""
    import summarymixing_b200 as S

    torch.manual_seed(13)
    m = _perturbed(S.SummaryMixing(256, 4, [256], 256, [256], 256, activation=S.Swish), 13).to(DEV).eval()
    B, T = 32, 1000
    x = torch.randn(B, T, 256, device=DEV)
    lens = torch.randint(300, T + 1, (B,))
    lens[0] = T
    mask = (torch.arange(T)[None] < lens[:, None]).to(DEV)
    dy = torch.randn(B, T, 256, device=DEV)

    def run(xx, mm, dd):
        for p in m.parameters():
            p.grad = None
        xx = xx.clone().requires_grad_(True)
        m(xx, src_padding_mask=mm).backward(dd)
        return xx.grad, {k: p.grad.clone() for k, p in m.named_parameters()}

    dx1, g1 = run(x, mask, dy)
    dx2, g2 = run(x, mask, 2.0 * dy)
    assert torch.isfinite(dx1).all()
    assert float((dx2 - 2.0 * dx1).abs().max()) <= 1e-5 * float(dx1.abs().max())
    for k in g1:
        assert float((g2[k] - 2.0 * g1[k]).abs().max()) <= 1e-4 * max(1e-6, float(g1[k].abs().max())), k
    assert float(dx1[~mask].abs().max()) == 0.0
    dx_solo, _ = run(x[5:6], mask[5:6], dy[5:6])
    assert float((dx_solo[0] - dx1[5]).abs().max()) <= 1e-5 * max(1.0, float(dx1[5].abs().max()))


def test_branchformer_layer_backward_at_recipe_dims_bf16_io():
    """One BranchformerEncoderLayer at the shipped recipe's dims (branchformer_summarymixing.yaml: D=512, csgu 3072, k=31,
    SummaryMixing-lite), bf16 activations, dropout 0: forward and gradients against torch.autograd of the oracle fed the same
    bf16-rounded input.  The gradients that travel between the layer's autograd nodes are bf16 like the activations (2^-8 relative
    rounding each): fp32 parameter gradients within 1e-2 of their scale (the lite cell's weight gradient is driven by B = 2 such
    rounded vectors), dx within 2^-6 of its scale (x feeds the residual and both branches: three rounded contributions, added and
    rounded again by autograd)."""
    import summarymixing_b200 as S

    torch.manual_seed(31)
    m = S.BranchformerEncoderLayer(512, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[512], local_proj_out_dim=512, summary_hid_dim=[512],
                                   summary_out_dim=512, mode="SummaryMixing-lite", dropout=0.0)
    m = _perturbed(m, 31)
    with torch.no_grad():
        m.convolution_branch.csgu.conv.conv.weight.add_(0.1 * torch.randn(m.convolution_branch.csgu.conv.conv.weight.shape))
    m = m.to(DEV).train()
    B, T = 2, 300
    x = torch.randn(B, T, 512, device=DEV).bfloat16().requires_grad_(True)
    mask = (torch.arange(T)[None] < torch.tensor([T, 171])[:, None]).to(DEV)
    dy = torch.randn(B, T, 512, device=DEV).bfloat16()
    y = m(x, src_key_padding_mask=mask)[0]
    assert y.dtype == torch.bfloat16 and y.requires_grad
    y.backward(dy)
    sd = {k: v.detach().cpu().float().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xo = x.detach().cpu().float().clone().requires_grad_(True)
    yo = O.branchformer_layer(xo, sd, "", act="gelu", mode="SummaryMixing-lite", src_key_padding_mask=mask.cpu())
    yo.backward(dy.detach().cpu().float())
    _close(y.float(), yo.detach(), 2 ** -7, "forward (bf16)")
    _close(x.grad.float(), xo.grad, 2 ** -6, "dx (bf16)")
    for k, p in m.named_parameters():
        if sd[k].grad is None:
            continue
        assert p.grad is not None and p.grad.dtype == torch.float32, k
        _close(p.grad, sd[k].grad, 1e-2, k)


def test_branchformer_backward_properties_and_sum_mask():
    """Utterance independence and padded-frame behaviour of the Branchformer layer's backward at D=256: an utterance's dx does not
    depend on the other utterances; with a (chunked) sum_mask forward and gradients follow the oracle."""
    import summarymixing_b200 as S

    torch.manual_seed(32)
    m = _perturbed(S.BranchformerEncoderLayer(256, 4, 31, csgu_linear_units=512, local_proj_hid_dim=[256], local_proj_out_dim=256,
                                              summary_hid_dim=[256], summary_out_dim=256, mode="SummaryMixing", dropout=0.0), 32).to(DEV).eval()
    B, T = 4, 200
    x = torch.randn(B, T, 256, device=DEV)
    mask = (torch.arange(T)[None] < torch.tensor([T, 150, 64, 199])[:, None]).to(DEV)
    dy = torch.randn(B, T, 256, device=DEV)
    xa = x.clone().requires_grad_(True)
    m(xa, src_key_padding_mask=mask)[0].backward(dy)
    xb = x[1:3].clone().requires_grad_(True)
    m(xb, src_key_padding_mask=mask[1:3])[0].backward(dy[1:3])
    _close(xa.grad[1:3], xb.grad.cpu(), 1e-5, "dx of a sub-batch")
    chunk = 32
    smask = (torch.arange(T)[None, :] // chunk <= torch.arange(T)[:, None] // chunk).float()   # a frame sees its own and earlier chunks
    m.zero_grad(set_to_none=True)   # (the two backward calls above accumulated into the parameters' .grad)
    xc = x.clone().requires_grad_(True)
    y = m(xc, src_mask=smask.to(DEV), src_key_padding_mask=mask)[0]
    y.backward(dy)
    sd = {k: v.detach().cpu().float().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xo = x.detach().cpu().clone().requires_grad_(True)
    yo = O.branchformer_layer(xo, sd, "", act="gelu", mode="SummaryMixing", src_mask=smask, src_key_padding_mask=mask.cpu())
    yo.backward(dy.cpu())
    _close(y, yo.detach(), 3e-4, "forward with sum_mask")
    _close(xc.grad, xo.grad, 3e-4, "dx with sum_mask")
    for k, p in m.named_parameters():
        if sd[k].grad is not None:
            _close(p.grad, sd[k].grad, 3e-4, k)


@pytest.mark.parametrize("structure", ["chunks", "weights"])
def test_cell_sum_mask_backward_both_forms(structure):
    """A chunked 0/1 sum mask takes the prefix-sum form of the summaries and of their gradient (O(T D) per utterance), a mask of
    arbitrary weights the (T,T) products: both against torch.autograd of the oracle, at a T that is not a multiple of the chunk."""
    import summarymixing_b200 as S

    torch.manual_seed(41)
    m = _perturbed(S.SummaryMixing(64, 4, [64], 64, [64], 64, activation=nn.GELU, global_dropout=0.0), 41).to(DEV).eval()
    B, T, chunk = 3, 203, 24
    x = torch.randn(B, T, 64, device=DEV, requires_grad=True)
    mask = (torch.arange(T)[None] < torch.tensor([T, 150, 31])[:, None]).to(DEV)
    dy = torch.randn(B, T, 64, device=DEV)
    if structure == "chunks":   # a frame sees the previous chunk, its own and nothing else (limited left context)
        ci = torch.arange(T) // chunk
        smask = ((ci[None, :] <= ci[:, None]) & (ci[None, :] >= ci[:, None] - 1)).float()
    else:
        smask = torch.rand(T, T, generator=torch.Generator().manual_seed(5)) + 0.1
    y = m(x, sum_mask=smask.to(DEV), src_padding_mask=mask)
    y.backward(dy)
    sd = {k: v.detach().cpu().float().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xo = x.detach().cpu().clone().requires_grad_(True)
    yo = O.summary_mixing(xo, sd, mode="SummaryMixing", act="gelu", use_layernorm=True, src_padding_mask=mask.cpu(), sum_mask=smask)
    yo.backward(dy.cpu())
    _close(y, yo.detach(), 1e-4, "forward")
    _close(x.grad, xo.grad, 1e-4, "dx")
    for k, p in m.named_parameters():
        _close(p.grad, sd[k].grad, 1e-4, k)
