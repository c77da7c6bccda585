"""autograd nodes over libsmx's backward entry points (smx_layernorm_bwd, smx_ffn_bwd, smx_conv_module_bwd; the cell's
node lives next to its module in nnet/summary_mixing.py).  Every backward call is self-contained — the library
recomputes what it needs from the saved input — so a node keeps only its input (and the mask).  Gradients come back in
fp32 and are cast to the parameter's dtype.  Training-mode dropout: a node called with p > 0 draws one seed from torch's
CPU generator (torch.manual_seed applies), runs smx_*_train_fwd and hands the same (p, seed) to smx_*_train_bwd, which
regenerates the counter-based masks (include/smx.h, smx_dropout)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _host as H
from . import _lib as L


def wants_grad(module, x) -> bool:
    """True when the call has to be recorded by autograd: x requires grad, or the module is in training mode with
    trainable parameters (eval-mode calls on plain inputs stay on the inference path)."""
    return torch.is_grad_enabled() and (x.requires_grad or (module.training and any(p.requires_grad for p in module.parameters())))


def new_dropout(module, p: float):
    """smx_dropout for one module call: None outside training mode or with p == 0 (dropout is then the identity)."""
    if not (module.training and p > 0):
        return None
    if p >= 1:
        raise ValueError("summarymixing_b200: dropout p must be < 1")
    return L.Dropout(float(p), int(torch.randint(0, 2 ** 62, (1,)).item()))


def same_p(*ps) -> float:
    """The one dropout probability a fused module call applies at all of its sites (the reference builds them from one argument)."""
    if any(p != ps[0] for p in ps):
        raise NotImplementedError(f"summarymixing_b200: one dropout probability per module call, got {ps}")
    return ps[0]


def pin_params(ctx, params) -> None:
    """Record the parameters' version counters at forward time.  The backward entry points recompute the forward from
    raw weight pointers (nothing but x is saved), so autograd cannot see a parameter that was modified in between; the
    node checks it itself and fails like torch does for a saved tensor modified in place."""
    ctx.param_versions = tuple(p._version for p in params)
    ctx.pinned_params = tuple(params)


def check_params(ctx) -> None:
    now = tuple(p._version for p in ctx.pinned_params)
    if now != ctx.param_versions:
        raise RuntimeError(
            "summarymixing_b200: a parameter of this module was modified in place between forward and backward "
            "(optimizer step or load_state_dict with the graph still alive); its backward recomputes the forward from "
            "the current weights and would return wrong gradients")


def _new_grads(params, dev):
    return [torch.empty(p.shape, dtype=torch.float32, device=dev) for p in params]


def _cast_out(ctx, first, grads, params):
    return [g.to(p.dtype) if ctx.needs_input_grad[first + i] else None for i, (g, p) in enumerate(zip(grads, params))]


class LayerNormFunction(torch.autograd.Function):
    """nn.LayerNorm over the last dim: smx_layernorm_fwd / smx_layernorm_bwd."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        from .nnet.containers import layer_norm

        ctx.eps = float(eps)
        ctx.save_for_backward(x, weight, bias)
        return layer_norm(x, weight, bias, eps)

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        dev = x.device
        xc, dyc = x.contiguous(), dy.contiguous()
        D = xc.shape[-1]
        rows = xc.numel() // D
        wv = H.WeightView()
        wv.stale((weight,), dev)
        dw, db = _new_grads((weight, bias), dev)
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            ws = H.workspace(dev, lib.smx_layernorm_bwd_workspace_bytes(dt, rows, D))
            L.check(lib.smx_layernorm_bwd(dt, rows, D, xc.data_ptr(), wv.ptr(weight, dev), ctx.eps, dyc.data_ptr(),
                                          H.p_or_none(dx), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(),
                                          H.stream_ptr(dev)))
        gw, gb = _cast_out(ctx, 1, (dw, db), (weight, bias))
        return dx, gw, gb, None


class FFNFunction(torch.autograd.Function):
    """y = x + 0.5 * FFN(LN(x)) (optionally followed by the layer's norm2): smx_ffn_fwd / smx_ffn_bwd.
    params = [ln.weight, ln.bias, W1, b1, W2, b2] (+ [norm.weight, norm.bias])."""

    @staticmethod
    def forward(ctx, fw, act, out_norm, drop, x, *params):
        # fw: L.FFNWeights (kept alive by the owning layer's WeightView); out_norm: (w_ptr, b_ptr, eps) or None; drop: L.Dropout or None
        dev = x.device
        xc = x.contiguous()
        rows = xc.numel() // xc.shape[-1]
        y = torch.empty_like(xc)
        lib = L.lib()
        dt = H.dtype_code(xc)
        ow, ob, oeps = out_norm if out_norm is not None else (None, None, 0.0)
        with torch.cuda.device(dev):
            if drop is None:
                ws = H.workspace(dev, lib.smx_ffn_workspace_bytes(C.byref(fw), dt, rows))
                L.check(lib.smx_ffn_fwd(C.byref(fw), act, dt, rows, xc.data_ptr(), ow, ob, oeps, y.data_ptr(), ws.data_ptr(),
                                        ws.numel(), H.stream_ptr(dev)))
            else:
                ws = H.workspace(dev, lib.smx_ffn_train_workspace_bytes(C.byref(fw), dt, rows, int(out_norm is not None)))
                L.check(lib.smx_ffn_train_fwd(C.byref(fw), act, dt, rows, xc.data_ptr(), ow, ob, oeps, C.byref(drop), y.data_ptr(),
                                              ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        ctx.drop = drop
        ctx.fw, ctx.act, ctx.out_norm, ctx.params = fw, act, out_norm, params  # fw carries its tensors (_keepalive)
        pin_params(ctx, params)
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        check_params(ctx)
        dev = xc.device
        dyc = dy.contiguous()
        rows = xc.numel() // xc.shape[-1]
        grads = _new_grads(ctx.params, dev)
        fg = L.FFNGrads()
        fg.ln_dw, fg.ln_db = grads[0].data_ptr(), grads[1].data_ptr()
        fg.w1.dw, fg.w1.db = grads[2].data_ptr(), grads[3].data_ptr()
        fg.w2.dw, fg.w2.db = grads[4].data_ptr(), grads[5].data_ptr()
        if ctx.out_norm is not None:
            fg.out_ln_dw, fg.out_ln_db = grads[6].data_ptr(), grads[7].data_ptr()
        ow, ob, oeps = ctx.out_norm if ctx.out_norm is not None else (None, None, 0.0)
        dx = torch.empty_like(xc) if ctx.needs_input_grad[4] else None
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            if ctx.drop is None:
                ws = H.workspace(dev, lib.smx_ffn_bwd_workspace_bytes(C.byref(ctx.fw), dt, rows, int(ctx.out_norm is not None)))
                L.check(lib.smx_ffn_bwd(C.byref(ctx.fw), ctx.act, dt, rows, xc.data_ptr(), ow, ob, oeps, dyc.data_ptr(),
                                        H.p_or_none(dx), C.byref(fg), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
            else:
                ws = H.workspace(dev, lib.smx_ffn_train_workspace_bytes(C.byref(ctx.fw), dt, rows, int(ctx.out_norm is not None)))
                L.check(lib.smx_ffn_train_bwd(C.byref(ctx.fw), ctx.act, dt, rows, xc.data_ptr(), ow, ob, oeps, C.byref(ctx.drop),
                                              dyc.data_ptr(), H.p_or_none(dx), C.byref(fg), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        return (None, None, None, None, dx, *_cast_out(ctx, 5, grads, ctx.params))


class ConvModuleFunction(torch.autograd.Function):
    """y = conv_module(x) * mask: smx_conv_module_fwd / smx_conv_module_bwd (chunk > 0: Dynamic Chunk Convolution,
    smx_conv_module_dcc_train_bwd).  params = [ln.weight, ln.bias, Wb, bb, dw.weight, dw.bias, after_ln.weight, after_ln.bias, Wo, bo]."""

    @staticmethod
    def forward(ctx, cw, act, drop, x, m8, chunk, *params):
        dev = x.device
        xc = x.contiguous()
        B, T, _ = xc.shape
        y = torch.empty_like(xc)
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            if drop is None:
                ws = H.workspace(dev, lib.smx_conv_module_workspace_bytes(C.byref(cw), dt, B, T))
                L.check(lib.smx_conv_module_fwd(C.byref(cw), act, dt, B, T, chunk, xc.data_ptr(), H.p_or_none(m8), None, y.data_ptr(),
                                                ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
            else:
                ws = H.workspace(dev, lib.smx_conv_module_train_workspace_bytes(C.byref(cw), dt, B, T))
                L.check(lib.smx_conv_module_dcc_train_fwd(C.byref(cw), act, dt, B, T, chunk, xc.data_ptr(), H.p_or_none(m8), C.byref(drop),
                                                          y.data_ptr(), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        ctx.drop, ctx.chunk = drop, chunk
        ctx.cw, ctx.act, ctx.params = cw, act, params
        pin_params(ctx, params)
        ctx.save_for_backward(xc, m8)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, m8 = ctx.saved_tensors
        check_params(ctx)
        dev = xc.device
        dyc = dy.contiguous()
        B, T, _ = xc.shape
        grads = _new_grads(ctx.params, dev)
        cg = L.ConvModGrads()
        cg.ln_dw, cg.ln_db = grads[0].data_ptr(), grads[1].data_ptr()
        cg.bottleneck.dw, cg.bottleneck.db = grads[2].data_ptr(), grads[3].data_ptr()
        cg.dw_dw, cg.dw_db = grads[4].data_ptr(), grads[5].data_ptr()
        cg.after_ln_dw, cg.after_ln_db = grads[6].data_ptr(), grads[7].data_ptr()
        cg.out.dw, cg.out.db = grads[8].data_ptr(), grads[9].data_ptr()
        dx = torch.empty_like(xc) if ctx.needs_input_grad[3] else None
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            ws = H.workspace(dev, lib.smx_conv_module_bwd_workspace_bytes(C.byref(ctx.cw), dt, B, T))
            if ctx.drop is None and ctx.chunk == 0:
                L.check(lib.smx_conv_module_bwd(C.byref(ctx.cw), ctx.act, dt, B, T, xc.data_ptr(), H.p_or_none(m8), dyc.data_ptr(),
                                                H.p_or_none(dx), C.byref(cg), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
            else:
                L.check(lib.smx_conv_module_dcc_train_bwd(C.byref(ctx.cw), ctx.act, dt, B, T, ctx.chunk, xc.data_ptr(), H.p_or_none(m8),
                                                          C.byref(ctx.drop) if ctx.drop is not None else None, dyc.data_ptr(),
                                                          H.p_or_none(dx), C.byref(cg), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        return (None, None, None, dx, None, None, *_cast_out(ctx, 6, grads, ctx.params))


class VanillaNNFunction(torch.autograd.Function):
    """n x (linear, act) (VanillaNN / ParallelLinear): smx_vanilla_nn_fwd / smx_vanilla_nn_bwd.  params = [W0, b0, W1, b1, ...]."""

    @staticmethod
    def forward(ctx, blocks, n, act, out_dim, x, *params):
        xc = x.contiguous()
        B, T = xc.shape[0], xc.shape[1]
        xc = xc.reshape(B, T, -1)
        dev = xc.device
        y = torch.empty(B, T, out_dim, dtype=xc.dtype, device=dev)
        lib = L.lib()
        dt = H.dtype_code(xc)
        rows = B * T
        with torch.cuda.device(dev):
            ws = H.workspace(dev, lib.smx_vanilla_nn_workspace_bytes(blocks, n, dt, rows))
            L.check(lib.smx_vanilla_nn_fwd(blocks, n, act, dt, rows, xc.data_ptr(), y.data_ptr(), ws.data_ptr(), ws.numel(),
                                           H.stream_ptr(dev)))
        ctx.blocks, ctx.n, ctx.act, ctx.params, ctx.x_shape = blocks, n, act, params, x.shape
        pin_params(ctx, params)
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        check_params(ctx)
        dev = xc.device
        dyc = dy.contiguous()
        rows = xc.shape[0] * xc.shape[1]
        grads = _new_grads(ctx.params, dev)
        lg = (L.LinearGrad * L.SMX_MAX_BLOCKS)()
        for i in range(ctx.n):
            lg[i].dw, lg[i].db = grads[2 * i].data_ptr(), grads[2 * i + 1].data_ptr()
        dx = torch.empty_like(xc) if ctx.needs_input_grad[4] else None
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            ws = H.workspace(dev, lib.smx_vanilla_nn_bwd_workspace_bytes(ctx.blocks, ctx.n, dt, rows))
            L.check(lib.smx_vanilla_nn_bwd(ctx.blocks, ctx.n, ctx.act, dt, rows, xc.data_ptr(), dyc.data_ptr(), H.p_or_none(dx), lg,
                                           ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        if dx is not None:
            dx = dx.reshape(ctx.x_shape)
        return (None, None, None, None, dx, *_cast_out(ctx, 5, grads, ctx.params))


class ConvBranchFunction(torch.autograd.Function):
    """ConvolutionBranch (Branchformer.py:86-97): smx_conv_branch_train_fwd / smx_conv_branch_train_bwd.
    params = [pre.weight, pre.bias, post.weight, post.bias, csgu.norm.weight, csgu.norm.bias, csgu.conv.weight, csgu.conv.bias]
    (+ [csgu.linear.weight, csgu.linear.bias])."""

    @staticmethod
    def forward(ctx, bw, drop, x, *params):
        dev = x.device
        xc = x.contiguous()
        B, T, _ = xc.shape
        y = torch.empty_like(xc)
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            ws = H.workspace(dev, lib.smx_conv_branch_train_workspace_bytes(C.byref(bw), dt, B, T))
            L.check(lib.smx_conv_branch_train_fwd(C.byref(bw), dt, B, T, xc.data_ptr(), C.byref(drop) if drop is not None else None,
                                                  y.data_ptr(), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        ctx.bw, ctx.drop, ctx.params = bw, drop, params
        pin_params(ctx, params)
        ctx.save_for_backward(xc)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        check_params(ctx)
        dev = xc.device
        dyc = dy.contiguous()
        B, T, _ = xc.shape
        grads = _new_grads(ctx.params, dev)
        bg = L.ConvBranchGrads()
        bg.pre.dw, bg.pre.db = grads[0].data_ptr(), grads[1].data_ptr()
        bg.post.dw, bg.post.db = grads[2].data_ptr(), grads[3].data_ptr()
        bg.csgu_ln_dw, bg.csgu_ln_db = grads[4].data_ptr(), grads[5].data_ptr()
        bg.csgu_dw_dw, bg.csgu_dw_db = grads[6].data_ptr(), grads[7].data_ptr()
        if len(grads) > 8:
            bg.csgu_linear.dw, bg.csgu_linear.db = grads[8].data_ptr(), grads[9].data_ptr()
        dx = torch.empty_like(xc) if ctx.needs_input_grad[2] else None
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            ws = H.workspace(dev, lib.smx_conv_branch_train_workspace_bytes(C.byref(ctx.bw), dt, B, T))
            L.check(lib.smx_conv_branch_train_bwd(C.byref(ctx.bw), dt, B, T, xc.data_ptr(),
                                                  C.byref(ctx.drop) if ctx.drop is not None else None, dyc.data_ptr(), H.p_or_none(dx),
                                                  C.byref(bg), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        return (None, None, dx, *_cast_out(ctx, 3, grads, ctx.params))


class DropoutFunction(torch.autograd.Function):
    """One nn.Dropout call of a layer's forward with libsmx's counter-based mask (smx_dropout_apply): y = x * keep(site) / (1 - p);
    the backward applies the same mask to dy."""

    @staticmethod
    def forward(ctx, drop, site, x):
        xc = x.contiguous()
        y = torch.empty_like(xc)
        dev = xc.device
        with torch.cuda.device(dev):
            L.check(L.lib().smx_dropout_apply(C.byref(drop), site, H.dtype_code(xc), xc.numel(), xc.data_ptr(), y.data_ptr(), H.stream_ptr(dev)))
        ctx.drop, ctx.site = drop, site
        return y

    @staticmethod
    def backward(ctx, dy):
        dyc = dy.contiguous()
        dx = torch.empty_like(dyc)
        dev = dyc.device
        with torch.cuda.device(dev):
            L.check(L.lib().smx_dropout_apply(C.byref(ctx.drop), ctx.site, H.dtype_code(dyc), dyc.numel(), dyc.data_ptr(), dx.data_ptr(),
                                              H.stream_ptr(dev)))
        return None, None, dx


def dropout(drop, site: int, x):
    """x when drop is None (eval mode or p == 0), else DropoutFunction."""
    return x if drop is None else DropoutFunction.apply(drop, site, x)
