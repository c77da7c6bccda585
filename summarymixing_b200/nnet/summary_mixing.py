"""SummaryMixing cell with the reference's surface (reference: speechbrain/nnet/summary_mixing.py).

Constructor signature, mode validation, sub-module names and state_dict keys follow summary_mixing.py:78-167;
``forward(x, sum_mask=None, src_padding_mask=None)`` follows :169-196.  The arithmetic runs in libsmx
(smx_summary_mixing_fwd).  Differences, deliberate: the default mask is built on x's device (the reference
allocates it on the CPU, :186, which breaks on GPU); dropout is the identity on the inference path and applied by the
training path (smx_summary_mixing_train_fwd / _bwd).
"""
from __future__ import annotations

from typing import Optional

import ctypes as C

import torch
import torch.nn as nn

from .. import _host as H
from .. import _lib as L
from ..lobes.models.VanillaNN import VanillaNN


class SummaryMixing(nn.Module):
    """SummaryMixing (https://arxiv.org/abs/2307.07421): y_t = combiner([f(x_t), mean_t s(x_t)]).

    Arguments are those of the reference class (summary_mixing.py:32-66).
    """

    def __init__(
        self,
        enc_dim,
        nhead,
        local_proj_hid_dim: Optional[list] = [512],
        local_proj_out_dim: Optional[int] = 512,
        summary_hid_dim: Optional[list] = [512],
        summary_out_dim: Optional[int] = 512,
        activation: Optional[nn.Module] = nn.GELU,
        global_dropout: Optional[float] = 0.1,
        mode: Optional[str] = "SummaryMixing",
        use_layernorm: Optional[bool] = True,
    ):
        super().__init__()
        if mode not in L.MODES:
            raise ValueError(
                "The SummaryMixing mode should either be 'SummaryMixing', 'SummaryMixing-lite', 'SummaryMixing-fast' or 'SummaryMixing-expdecay'"
            )
        self.enc_dim, self.nhead, self.mode = enc_dim, nhead, mode
        self.local_proj_hid_dim, self.local_proj_out_dim = local_proj_hid_dim, local_proj_out_dim
        self.summary_hid_dim, self.summary_out_dim = summary_hid_dim, summary_out_dim
        self.use_layernorm = use_layernorm
        self.activation = activation()
        self.dropout = nn.Dropout(global_dropout)

        def mlp(in_dim, widths, heads=1):
            # VanillaNN over the last dim of a (B,T,in_dim) input: one (linear, activation) pair per width
            return VanillaNN(input_shape=[None, None, in_dim], dnn_blocks=len(widths), dnn_neurons=list(widths),
                             activation=activation, n_split=heads)

        # Which sub-networks a mode owns (summary_mixing.py:103-161).  Registration order = the reference's, so that
        # state_dict() and parameters() enumerate identically:
        #   full / expdecay: local_proj (f), summary_local_merging (combiner over [f ; mean s]), summary_proj (s)
        #   lite:            summary_proj only
        #   fast:            global_proj (one dense D -> 2 D_l layer whose halves are f and s), summary_local_merging
        if mode in ("SummaryMixing", "SummaryMixing-expdecay"):
            self.local_proj = mlp(enc_dim, local_proj_hid_dim + [local_proj_out_dim], nhead)
            self.summary_local_merging = mlp(local_proj_out_dim + summary_out_dim, [summary_out_dim])
        if mode == "SummaryMixing-fast":
            self.global_proj = mlp(enc_dim, [2 * local_proj_out_dim])
            self.summary_local_merging = mlp(2 * local_proj_out_dim, [summary_out_dim])
        else:
            self.summary_proj = mlp(enc_dim, summary_hid_dim + [summary_out_dim], nhead)
        if mode == "SummaryMixing-expdecay":
            self.decay_constant = nn.Parameter(torch.tensor(0.995), requires_grad=False)  # frozen (:158-161)
        if use_layernorm:
            # owned in every mode, like the reference (:163-165), although lite and fast never apply them
            self.local_norm = nn.LayerNorm(local_proj_out_dim)
            self.summary_norm = nn.LayerNorm(summary_out_dim)

        self.apply(self._init_parameters)  # zero the biases of dense layers (:167, 326-328)
        self._act_code = H.act_code(self.activation)
        self._wv = H.WeightView()

    def _init_parameters(self, module):
        if isinstance(module, nn.Linear):
            torch.nn.init.zeros_(module.bias)

    # -- library plumbing -------------------------------------------------------------------------
    def params(self):
        return [p for p in self.parameters()]

    def fill(self, cw: L.CellWeights, wv: H.WeightView, device) -> None:
        """Fill an smx_cell_weights from this module's parameters."""
        cw.mode = L.MODES[self.mode]
        cw.act = self._act_code
        cw.use_layernorm = int(bool(self.use_layernorm))
        cw.enc_dim = self.enc_dim
        cw.local_out_dim = self.local_proj_out_dim
        cw.summary_out_dim = self.summary_out_dim
        cw.n_local = cw.n_summary = 0
        if hasattr(self, "local_proj"):
            cw.n_local = self.local_proj.fill(cw.local, wv, device)
        if hasattr(self, "summary_proj"):
            cw.n_summary = self.summary_proj.fill(cw.summary, wv, device)
        if hasattr(self, "global_proj"):
            gp = (L.Linear * L.SMX_MAX_BLOCKS)()
            self.global_proj.fill(gp, wv, device)
            cw.global_proj = gp[0]
        if hasattr(self, "summary_local_merging"):
            mg = (L.Linear * L.SMX_MAX_BLOCKS)()
            self.summary_local_merging.fill(mg, wv, device)
            cw.merge = mg[0]
        if self.use_layernorm:
            cw.local_norm_w = wv.ptr(self.local_norm.weight, device)
            cw.local_norm_b = wv.ptr(self.local_norm.bias, device)
            cw.summary_norm_w = wv.ptr(self.summary_norm.weight, device)
            cw.summary_norm_b = wv.ptr(self.summary_norm.bias, device)
        if self.mode == "SummaryMixing-expdecay":
            cw.decay_constant = float(self.decay_constant)
        H.pack_tc(wv, device, cw, L.lib().smx_cell_packed_bytes, L.lib().smx_cell_pack)

    @property
    def out_dim(self) -> int:
        return self.summary_out_dim

    def forward(self, x, sum_mask=None, src_padding_mask=None):
        """x: (B,T,enc_dim); sum_mask: (T,T) or None; src_padding_mask: (B,T), 1/True = valid frame.
        Returns (B,T,summary_out_dim) in x's dtype (lite: a stride-0 expand over T, as the reference, :322).

        Differentiable (smx_summary_mixing_bwd and its training-mode / sum_mask forms) in all four modes when x requires grad or in
        training mode (dropout on the concatenation applied; "-expdecay": decay_constant is not trainable, as in the reference)."""
        H.require_cuda(x, "SummaryMixing")
        if x.dim() != 3 or x.shape[-1] != self.enc_dim:
            raise RuntimeError(f"SummaryMixing expects (B,T,{self.enc_dim}), got {tuple(x.shape)}")
        B, T, _ = x.shape
        dev = x.device
        mask = H.mask_u8(src_padding_mask, B, T, dev)
        smask = H.sum_mask_f32(sum_mask, T, dev)
        if torch.is_grad_enabled() and (x.requires_grad or (self.training and any(p.requires_grad for p in self.parameters()))):
            lite = self.mode == "SummaryMixing-lite"  # (lite ignores sum_mask, summary_mixing.py:300-324)
            from .. import _autograd as A

            drop = None if lite else A.new_dropout(self, self.dropout.p)  # (lite has no dropout, summary_mixing.py:300-324)
            y = _CellFunction.apply(self, x, mask, None if lite else smask, drop, *self.grad_params())
            return y.unsqueeze(1).expand(-1, T, -1) if lite else y
        H.check_grad_mode(self)
        return self._forward_impl(x, mask, smask)

    def grad_params(self):
        """Parameters in the order smx_cell_grads lists their gradients (mode "SummaryMixing")."""
        if self.mode == "SummaryMixing-lite":
            return self.summary_proj.params()
        if self.mode == "SummaryMixing-fast":
            return self.global_proj.params() + self.summary_local_merging.params()
        out = self.local_proj.params() + self.summary_proj.params() + self.summary_local_merging.params()
        if self.use_layernorm:
            out += [self.local_norm.weight, self.local_norm.bias, self.summary_norm.weight, self.summary_norm.bias]
        return out

    def _weights(self, dev):
        if self._wv.stale(self.params(), dev):
            cw = L.CellWeights()
            self.fill(cw, self._wv, dev)
            self._wv.struct = cw
        return self._wv.struct

    def _forward_impl(self, x, mask, smask, expand_lite=True):
        B, T, _ = x.shape
        dev = x.device
        xc = x.contiguous()
        cw = self._weights(dev)
        lite = self.mode == "SummaryMixing-lite"
        y = torch.empty((B, self.summary_out_dim) if lite else (B, T, self.summary_out_dim), dtype=x.dtype, device=dev)
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            nbytes = lib.smx_summary_mixing_workspace_bytes(cw, dt, B, T, int(smask is not None))
            ws = H.workspace(dev, nbytes)
            L.check(lib.smx_summary_mixing_fwd(cw, dt, B, T, xc.data_ptr(), H.p_or_none(mask),
                                               H.p_or_none(smask), None, y.data_ptr(), ws.data_ptr(), ws.numel(),
                                               H.stream_ptr(dev)))
        if lite and expand_lite:
            return y.unsqueeze(1).expand(-1, T, -1)
        return y

    def _train_forward_impl(self, x, mask, drop, smask=None):
        """Training-mode forward with dropout on the concatenation: smx_summary_mixing_train_fwd (sum_mask: the masked form)."""
        B, T, _ = x.shape
        dev = x.device
        xc = x.contiguous()
        cw = self._weights(dev)
        y = torch.empty((B, T, self.summary_out_dim), dtype=x.dtype, device=dev)
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            if smask is None:
                ws = H.workspace(dev, lib.smx_summary_mixing_train_workspace_bytes(cw, dt, B, T))
                L.check(lib.smx_summary_mixing_train_fwd(cw, dt, B, T, xc.data_ptr(), H.p_or_none(mask), C.byref(drop), y.data_ptr(),
                                                         ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
            else:
                ws = H.workspace(dev, lib.smx_summary_mixing_masked_train_workspace_bytes(cw, dt, B, T))
                L.check(lib.smx_summary_mixing_masked_train_fwd(cw, dt, B, T, xc.data_ptr(), H.p_or_none(mask), smask.data_ptr(), C.byref(drop),
                                                                y.data_ptr(), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        return y

    def _backward_impl(self, x, mask, dy, want_dx, cw=None, drop=None, smask=None):
        """(dx or None, [fp32 gradient per grad_params() entry]) through smx_summary_mixing_bwd."""
        B, T, _ = x.shape
        dev = x.device
        xc, dyc = x.contiguous(), dy.contiguous()
        if cw is None:
            cw = self._weights(dev)
        plist = self.grad_params()
        grads = [torch.empty(p.shape, dtype=torch.float32, device=dev) for p in plist]
        cg = L.CellGrads()
        it = iter(grads)
        lite, fast = self.mode == "SummaryMixing-lite", self.mode == "SummaryMixing-fast"
        if fast:
            nets = ()
            cg.global_proj.dw = next(it).data_ptr()
            cg.global_proj.db = next(it).data_ptr()
        elif lite:
            nets = ((cg.summary, self.summary_proj),)
        else:
            nets = ((cg.local, self.local_proj), (cg.summary, self.summary_proj))
        for dst, net in nets:
            for i in range(len(net._linears)):
                dst[i].dw = next(it).data_ptr()
                dst[i].db = next(it).data_ptr()
        if not lite:
            cg.merge.dw = next(it).data_ptr()
            cg.merge.db = next(it).data_ptr()
        if self.use_layernorm and not lite and not fast:
            cg.local_norm_dw, cg.local_norm_db = next(it).data_ptr(), next(it).data_ptr()
            cg.summary_norm_dw, cg.summary_norm_db = next(it).data_ptr(), next(it).data_ptr()
        dx = torch.empty_like(xc) if want_dx else None
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            if smask is not None:
                ws = H.workspace(dev, lib.smx_summary_mixing_masked_train_workspace_bytes(cw, dt, B, T))
                L.check(lib.smx_summary_mixing_masked_train_bwd(cw, dt, B, T, xc.data_ptr(), H.p_or_none(mask), smask.data_ptr(),
                                                                C.byref(drop) if drop is not None else None, dyc.data_ptr(), H.p_or_none(dx),
                                                                C.byref(cg), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
            elif drop is None:
                nbytes = lib.smx_summary_mixing_bwd_workspace_bytes(cw, dt, B, T)
                ws = H.workspace(dev, nbytes)
                L.check(lib.smx_summary_mixing_bwd(cw, dt, B, T, xc.data_ptr(), H.p_or_none(mask), dyc.data_ptr(),
                                                   H.p_or_none(dx), C.byref(cg), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
            else:
                ws = H.workspace(dev, lib.smx_summary_mixing_train_workspace_bytes(cw, dt, B, T))
                L.check(lib.smx_summary_mixing_train_bwd(cw, dt, B, T, xc.data_ptr(), H.p_or_none(mask), C.byref(drop), dyc.data_ptr(),
                                                         H.p_or_none(dx), C.byref(cg), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        return dx, grads


class _CellFunction(torch.autograd.Function):
    """autograd node of the cell: forward = smx_summary_mixing_fwd, backward = smx_summary_mixing_bwd (which recomputes
    the intermediates from x, so only x and the mask are kept)."""

    @staticmethod
    def forward(ctx, module, x, mask, smask, drop, *params):
        from .. import _autograd as A

        ctx.module = module
        ctx.drop = drop
        ctx.save_for_backward(x, mask, smask)
        y = module._forward_impl(x, mask, smask, expand_lite=False) if drop is None else module._train_forward_impl(x, mask, drop, smask)
        ctx.cw = module._wv.struct  # the weight struct of THIS forward (with the tensors it points into)
        A.pin_params(ctx, module.params())
        return y

    @staticmethod
    def backward(ctx, dy):
        from .. import _autograd as A

        x, mask, smask = ctx.saved_tensors
        A.check_params(ctx)
        module = ctx.module
        dx, grads = module._backward_impl(x, mask, dy, ctx.needs_input_grad[1], cw=ctx.cw, drop=ctx.drop, smask=smask)
        plist = module.grad_params()
        out = [g.to(p.dtype) if ctx.needs_input_grad[5 + i] else None for i, (g, p) in enumerate(zip(grads, plist))]
        return (None, dx, None, None, None, *out)
