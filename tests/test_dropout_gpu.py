"""GPU: training-mode dropout (smx_*_train_fwd / smx_*_train_bwd through the modules' autograd nodes).  torch.nn.Dropout's random
stream cannot be matched; the parity statement is: with libsmx's counter-based masks (restated bit-exactly in oracle/dropout.py)
in place of torch's, forward and every gradient equal the reference algorithm's (the oracle under torch.autograd).
Tolerances as tests/test_backward_gpu.py (fp32 I/O): <= 3e-4 x max(1, |reference|max)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn as nn

import summarymixing_b200 as S
from oracle import dropout as OD
from oracle import smx_oracle as O
from summarymixing_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, rel, what):
    err = float((a.detach().cpu().double() - b.double()).abs().max())
    assert err <= rel * max(1.0, float(b.abs().max())), f"{what}: max-abs {err:.3e}"


def _seeds(seed, n):
    """The seeds the next n module calls draw (A.new_dropout: one torch.randint(0, 2**62) from the CPU generator per call)."""
    torch.manual_seed(seed)
    out = [int(torch.randint(0, 2 ** 62, (1,)).item()) for _ in range(n)]
    torch.manual_seed(seed)
    return out


def _perturbed(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
    return m


def _oracle(fn, m, x, dy):
    sd = {k: v.detach().cpu().float().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xo = x.detach().cpu().float().clone().requires_grad_(True)
    y = fn(xo, sd)
    y.backward(dy.detach().cpu().float())
    return y.detach(), xo.grad, {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("p,seed,site,n", [(0.1, 1, 0, 1000), (0.5, 2 ** 61 + 12345, 1, 4097), (0.25, 987654321987, 0, 3), (0.9, 7, 3, (1 << 21) + 5)])
def test_keep_mask_matches_the_oracle_bit_for_bit(p, seed, site, n):
    keep = torch.empty(n, dtype=torch.uint8, device=DEV)
    d = L.Dropout(p, seed)
    L.check(L.lib().smx_dropout_keep_mask(C.byref(d), site, n, keep.data_ptr(), None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(keep.cpu().numpy(), OD.keep_mask(p, seed, site, n))


@pytest.mark.parametrize("mode,use_ln", [("SummaryMixing", True), ("SummaryMixing", False), ("SummaryMixing-fast", True)])
def test_cell_dropout_forward_and_gradients(mode, use_ln):
    torch.manual_seed(3)
    p = 0.2
    m = _perturbed(S.SummaryMixing(64, 4, [64], 64, [64], 64, activation=nn.GELU, global_dropout=p, mode=mode, use_layernorm=use_ln), 3)
    m = m.to(DEV).train()
    B, T = 3, 77
    x = torch.randn(B, T, 64, device=DEV, requires_grad=True)
    mask = (torch.arange(T)[None] < torch.tensor([T, 40, 5])[:, None]).to(DEV)
    dy = torch.randn(B, T, 64, device=DEV)
    (s0,) = _seeds(41, 1)
    y = m(x, src_padding_mask=mask)
    y.backward(dy)
    hook = OD.Hook(p, {"cat": (s0, 0)})
    y_or, dx_or, g_or = _oracle(lambda xo, sd: O.summary_mixing(xo, sd, mode=mode, act="gelu", use_layernorm=use_ln,
                                                                src_padding_mask=mask.cpu(), drop=hook), m, x, dy)
    assert hook.used == ["cat"]
    _close(y, y_or, 1e-4, "forward")
    _close(x.grad, dx_or, 1e-4, "dx")
    for k, prm in m.named_parameters():
        if g_or[k] is None:  # a parameter this mode does not use (the "-fast" cell keeps the unused projections)
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, k
            continue
        _close(prm.grad, g_or[k], 1e-4, k)
    # a second call draws another seed: another mask; eval mode: no dropout
    y2 = m(x.detach(), src_padding_mask=mask)
    assert float((y2.detach() - y.detach()).abs().max()) > 1e-3
    with torch.no_grad():
        y_eval = m.eval()(x.detach(), src_padding_mask=mask)
    y_plain = O.summary_mixing(x.detach().cpu(), dict(m.cpu().state_dict()), mode=mode, act="gelu", use_layernorm=use_ln,
                               src_padding_mask=mask.cpu())
    _close(y_eval, y_plain, 1e-4, "eval forward")


def test_conv_module_dropout_forward_and_gradients():
    torch.manual_seed(4)
    p = 0.15
    m = _perturbed(S.ConvolutionModule(64, 15, dropout=p, masked_false_or_true=False), 4).to(DEV).train()
    B, T = 2, 90
    x = torch.randn(B, T, 64, device=DEV, requires_grad=True)
    mask = (torch.arange(T)[None] < torch.tensor([T, 33])[:, None]).to(DEV)
    dy = torch.randn(B, T, 64, device=DEV)
    (s0,) = _seeds(42, 1)
    y = m(x, mask.unsqueeze(-1))
    y.backward(dy)
    hook = OD.Hook(p, {"out": (s0, 0)})
    y_or, dx_or, g_or = _oracle(lambda xo, sd: O.convolution_module(xo, sd, "", act="swish", mask=mask.cpu().unsqueeze(-1), drop=hook), m, x, dy)
    _close(y, y_or, 1e-4, "forward")
    _close(x.grad, dx_or, 2e-4, "dx")
    for k, prm in m.named_parameters():
        _close(prm.grad, g_or[k], 2e-4, k)


def _layer_sites(prefix, s):
    # draw order of ConformerEncoderLayer._forward_autograd: ffn_module1, ffn_module2, the cell, the convolution module
    return {prefix + "ffn_module1.inner": (s[0], 0), prefix + "ffn_module1.outer": (s[0], 1),
            prefix + "ffn_module2.inner": (s[1], 0), prefix + "ffn_module2.outer": (s[1], 1),
            prefix + "mha_layer.cat": (s[2], 0), prefix + "convolution_module.out": (s[3], 0)}


def test_encoder_trains_with_the_recipes_dropout():
    """Two layers at dropout 0.1 (the shipped recipes' value, conformer_summarymixing.yaml): forward and every gradient against the
    oracle with the same masks; gradients reach every parameter; the masks of two steps differ."""
    torch.manual_seed(5)
    p, nl = 0.1, 2
    enc = _perturbed(S.ConformerEncoder(nl, 64, 128, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[64], local_proj_out_dim=64,
                                        summary_hid_dim=[64], dropout=p), 5).to(DEV).train()
    B, T = 3, 90
    x = torch.randn(B, T, 64, device=DEV, requires_grad=True)
    mask = (torch.arange(T)[None] < torch.tensor([T, 50, 17])[:, None]).to(DEV)
    dy = torch.randn(B, T, 64, device=DEV)
    s = _seeds(43, 4 * nl)
    y = enc(x, src_key_padding_mask=mask)[0]
    y.backward(dy)
    sites = {}
    for i in range(nl):
        sites.update(_layer_sites(f"layers.{i}.", s[4 * i:4 * i + 4]))
    hook = OD.Hook(p, sites)
    y_or, dx_or, g_or = _oracle(lambda xo, sd: O.conformer_encoder(xo, sd, nl, act="swish", src_key_padding_mask=mask.cpu(), drop=hook), enc, x, dy)
    assert sorted(hook.used) == sorted(sites)
    _close(y, y_or, 5e-4, "forward")
    _close(x.grad, dx_or, 5e-4, "dx")
    for k, prm in enc.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad).all(), k
        _close(prm.grad, g_or[k], 5e-4, k)
    y2 = enc(x.detach(), src_key_padding_mask=mask)[0]
    assert float((y2.detach() - y.detach()).abs().max()) > 1e-3


def test_dropout_bf16_io_and_keep_rate():
    """bf16 activations on the training path, and the empirical keep rate of the cell's mask at the bench width."""
    torch.manual_seed(6)
    p = 0.1
    m = S.ConformerEncoderLayer(256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256], local_proj_out_dim=256,
                                summary_hid_dim=[256], dropout=p).to(DEV).train()
    B, T = 4, 300
    x = torch.randn(B, T, 256, device=DEV).bfloat16().requires_grad_(True)
    mask = torch.ones(B, T, dtype=torch.bool, device=DEV)
    s = _seeds(44, 4)
    y = m(x, src_key_padding_mask=mask)[0]
    y.float().pow(2).mean().backward()
    assert y.dtype == torch.bfloat16 and x.grad.dtype == torch.bfloat16 and torch.isfinite(x.grad.float()).all()
    hook = OD.Hook(p, _layer_sites("", s))
    y_or = O.conformer_layer(x.detach().float().cpu(), {k: v.float().cpu() for k, v in m.state_dict().items()}, "", act="swish",
                             src_key_padding_mask=mask.cpu(), drop=hook)
    rel = float((y.detach().float().cpu() - y_or).norm() / y_or.norm())
    assert rel < 1e-2, rel  # the result is rounded to bf16 once; the arithmetic is fp32
    keep = torch.empty(B * T * 512, dtype=torch.uint8, device=DEV)
    L.check(L.lib().smx_dropout_keep_mask(C.byref(L.Dropout(p, s[2])), 0, keep.numel(), keep.data_ptr(), None))
    assert abs(float(keep.float().mean()) - (1 - p)) < 2e-3


def _csgu_live(enc, seed):
    """The CSGU's depthwise weights are ~1e-6 at initialisation (upstream): give the convolution something to do."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, prm in enc.named_parameters():
            if "csgu.conv.conv.weight" in n:
                prm.add_(0.2 * torch.randn(prm.shape, generator=g))
    return enc


@pytest.mark.parametrize("mode", ["SummaryMixing", "SummaryMixing-lite"])
def test_branchformer_encoder_trains_with_dropout(mode):
    """Two Branchformer layers at the recipes' dropout 0.1 (branchformer_summarymixing.yaml): forward and every gradient against the
    oracle under the same masks.  Seeds are drawn per layer call in the order layer (its three nn.Dropout calls are sites 1, 2, 3),
    cell (mode "SummaryMixing" only: the reference builds the cell with its default global_dropout = 0.1), convolution branch."""
    torch.manual_seed(8)
    p, nl, lite = 0.1, 2, mode == "SummaryMixing-lite"
    enc = S.BranchformerEncoder(nl, 64, 1, 31, csgu_linear_units=192, local_proj_hid_dim=[64], local_proj_out_dim=64, summary_hid_dim=[64],
                                summary_out_dim=64, mode=mode, dropout=p)
    enc = _csgu_live(_perturbed(enc, 8), 9).to(DEV).train()
    B, T = 3, 70
    x = torch.randn(B, T, 64, device=DEV, requires_grad=True)
    mask = (torch.arange(T)[None] < torch.tensor([T, 41, 16])[:, None]).to(DEV)
    dy = torch.randn(B, T, 64, device=DEV)
    per = 2 if lite else 3
    s = _seeds(45, per * nl)
    y = enc(x, src_key_padding_mask=mask)[0]
    y.backward(dy)
    sites = {}
    for i in range(nl):
        pre, si = f"layers.{i}.", s[per * i:per * (i + 1)]
        sites.update({pre + "x1": (si[0], 1), pre + "x2": (si[0], 2), pre + "merge": (si[0], 3),
                      pre + "convolution_branch.csgu.out": (si[-1], 0)})
        if not lite:
            sites[pre + "mha_layer.cat"] = (si[1], 0)
    hook = OD.Hook(p, sites)
    y_or, dx_or, g_or = _oracle(lambda xo, sd: O.branchformer_encoder(xo, sd, nl, act="gelu", mode=mode, src_key_padding_mask=mask.cpu(),
                                                                      drop=hook), enc, x, dy)
    assert sorted(hook.used) == sorted(sites)
    _close(y, y_or, 5e-4, "forward")
    _close(x.grad, dx_or, 5e-4, "dx")
    for k, prm in enc.named_parameters():
        if g_or[k] is None:  # "-lite" keeps the unused projections of the cell
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, k
            continue
        assert prm.grad is not None and torch.isfinite(prm.grad).all(), k
        _close(prm.grad, g_or[k], 5e-4, k)
    y2 = enc(x.detach(), src_key_padding_mask=mask)[0]
    assert float((y2.detach() - y.detach()).abs().max()) > 1e-3


def test_convolution_branch_standalone_and_use_linear_after_conv():
    """ConvolutionBranch.forward on its own (Branchformer.py:86-97) with the optional linear after the convolution and a gate
    activation: eval forward, and forward + gradients in training mode with the CSGU's dropout, against the oracle."""
    torch.manual_seed(10)
    p = 0.2
    m = S.ConvolutionBranch(64, linear_units=128, kernel_size=15, activation=nn.GELU, gate_activation=nn.Tanh, dropout=p, use_linear_after_conv=True)
    m = _csgu_live(_perturbed(m, 10), 11).to(DEV)
    B, T = 2, 45
    x = torch.randn(B, T, 64, device=DEV, requires_grad=True)
    dy = torch.randn(B, T, 64, device=DEV)
    with torch.no_grad():
        y_eval = m.eval()(x.detach())
    y_plain = O.convolution_branch(x.detach().cpu(), {k: v.detach().cpu() for k, v in m.state_dict().items()}, "", act="gelu", gate_act="tanh")
    _close(y_eval, y_plain, 1e-4, "eval forward")
    m.train()
    (s0,) = _seeds(46, 1)
    y = m(x)
    y.backward(dy)
    hook = OD.Hook(p, {"csgu.out": (s0, 0)})
    y_or, dx_or, g_or = _oracle(lambda xo, sd: O.convolution_branch(xo, sd, "", act="gelu", gate_act="tanh", drop=hook), m, x, dy)
    assert hook.used == ["csgu.out"]
    _close(y, y_or, 1e-4, "forward")
    _close(x.grad, dx_or, 2e-4, "dx")
    for k, prm in m.named_parameters():
        _close(prm.grad, g_or[k], 2e-4, k)
    # T too short for the reflect padding: a loud error like torch's
    with pytest.raises(Exception):
        m(torch.randn(1, 7, 64, device=DEV))


@pytest.mark.parametrize("mode", ["SummaryMixing", "SummaryMixing-fast", "SummaryMixing-expdecay"])
def test_cell_dropout_with_sum_mask(mode):
    """Dynamic Chunk Training with dropout: per-frame summaries under a chunked sum_mask, dropout on the written-out concatenation
    (smx_summary_mixing_masked_train_fwd / _bwd) against the oracle under the same mask; "-expdecay": the Laplace weights times the
    chunk mask (summary_mixing.py:223-224)."""
    torch.manual_seed(13)
    p = 0.2
    m = _perturbed(S.SummaryMixing(64, 4, [64], 64, [64], 64, activation=nn.GELU, global_dropout=p, mode=mode, use_layernorm=(mode != "SummaryMixing-fast")), 13)
    m = m.to(DEV).train()
    B, T, chunk = 3, 77, 16
    x = torch.randn(B, T, 64, device=DEV, requires_grad=True)
    mask = (torch.arange(T)[None] < torch.tensor([T, 40, 21])[:, None]).to(DEV)
    smask = (torch.arange(T)[None, :] // chunk <= torch.arange(T)[:, None] // chunk).float()
    dy = torch.randn(B, T, 64, device=DEV)
    (s0,) = _seeds(47, 1)
    y = m(x, sum_mask=smask.to(DEV), src_padding_mask=mask)
    y.backward(dy)
    hook = OD.Hook(p, {"cat": (s0, 0)})
    y_or, dx_or, g_or = _oracle(lambda xo, sd: O.summary_mixing(xo, sd, mode=mode, act="gelu", use_layernorm=(mode != "SummaryMixing-fast"),
                                                                src_padding_mask=mask.cpu(), sum_mask=smask, drop=hook), m, x, dy)
    assert hook.used == ["cat"]
    _close(y, y_or, 1e-4, "forward")
    _close(x.grad, dx_or, 1e-4, "dx")
    for k, prm in m.named_parameters():
        if not prm.requires_grad:  # decay_constant of "-expdecay" (requires_grad=False in the reference, summary_mixing.py:159-161)
            continue
        if g_or[k] is None:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, k
            continue
        _close(prm.grad, g_or[k], 1e-4, k)
