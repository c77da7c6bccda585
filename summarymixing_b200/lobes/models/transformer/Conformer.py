"""Conformer encoder blocks around the SummaryMixing cell, with the reference's surface
(reference: speechbrain/lobes/models/transformer/Conformer.py).

ConvolutionModule (:80-340), ConformerEncoderLayer (:343-548) and ConformerEncoder (:652-827) keep the
reference constructor signatures, forward contracts and state_dict keys.  Only attention_type ==
"SummaryMixing" is built (this package is the SummaryMixing hot path); the decoder and the streaming
entry points (:550-649, :829-1192) are out of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .... import _autograd as A
from .... import _host as H
from .... import _lib as L
from ....nnet.activations import Swish
from ....nnet.containers import LayerNorm, PositionalwiseFeedForward
from ....nnet.summary_mixing import SummaryMixing


def _chunk_size(dynchunktrain_config) -> int:
    if dynchunktrain_config is None:
        return 0
    return int(dynchunktrain_config.chunk_size)


class ConvolutionModule(nn.Module):
    """LN -> pointwise Conv1d(D,2D)+GLU -> depthwise Conv1d(k) -> LN -> act -> Linear -> (dropout) -> mask.

    Arguments as the reference (Conformer.py:83-100).  ``masked_false_or_true=False`` (the SummaryMixing
    convention, :334-338) multiplies the output by the (B,T,1) mask; True would masked_fill where the
    mask is True.
    """

    def __init__(
        self,
        input_size,
        kernel_size=31,
        bias=True,
        activation=Swish,
        dropout=0.0,
        causal=False,
        dilation=1,
        masked_false_or_true=True,
    ):
        super().__init__()
        if dilation != 1:
            raise NotImplementedError("libsmx implements dilation == 1 (no shipped recipe uses another value)")
        if not bias:
            raise NotImplementedError("libsmx implements bias=True convolution modules")
        self.kernel_size = kernel_size
        self.causal = causal
        self.dilation = dilation
        self.masked_false_or_true = masked_false_or_true
        if self.causal:
            self.padding = (kernel_size - 1) * 2 ** (dilation - 1)
        else:
            self.padding = (kernel_size - 1) * 2 ** (dilation - 1) // 2

        self.layer_norm = nn.LayerNorm(input_size)
        self.bottleneck = nn.Sequential(
            nn.Conv1d(input_size, 2 * input_size, kernel_size=1, stride=1, bias=bias), nn.GLU(dim=1),
        )
        self.conv = nn.Conv1d(
            input_size, input_size, kernel_size=kernel_size, stride=1, padding=self.padding, dilation=dilation,
            groups=input_size, bias=bias,
        )
        self.after_conv = nn.Sequential(
            nn.LayerNorm(input_size), activation(), nn.Linear(input_size, input_size, bias=bias), nn.Dropout(dropout),
        )
        self.input_size = input_size
        self._act_code = H.act_code(self.after_conv[1])
        self._wv = H.WeightView()

    def params(self):
        return list(self.parameters())

    def grad_params(self):
        """Parameters in the order smx_convmod_grads lists their gradients."""
        return [self.layer_norm.weight, self.layer_norm.bias, self.bottleneck[0].weight, self.bottleneck[0].bias,
                self.conv.weight, self.conv.bias, self.after_conv[0].weight, self.after_conv[0].bias,
                self.after_conv[2].weight, self.after_conv[2].bias]

    def fill(self, cw: L.ConvModWeights, wv: H.WeightView, device) -> None:
        D = self.input_size
        cw.ln_w = wv.ptr(self.layer_norm.weight, device)
        cw.ln_b = wv.ptr(self.layer_norm.bias, device)
        H.fill_linear(cw.bottleneck, wv, device, self.bottleneck[0].weight, self.bottleneck[0].bias, D, 2 * D)
        cw.dw_w = wv.ptr(self.conv.weight, device)
        cw.dw_b = wv.ptr(self.conv.bias, device)
        cw.after_ln_w = wv.ptr(self.after_conv[0].weight, device)
        cw.after_ln_b = wv.ptr(self.after_conv[0].bias, device)
        H.fill_linear(cw.out, wv, device, self.after_conv[2].weight, self.after_conv[2].bias, D, D)
        cw.kernel_size = self.kernel_size
        cw.causal = int(self.causal)
        H.pack_tc(wv, device, cw, L.lib().smx_convmod_packed_bytes, L.lib().smx_convmod_pack)

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, dynchunktrain_config=None):
        """x: (B,T,D); mask: (B,T,1) or (B,T) in the convention selected by ``masked_false_or_true``."""
        H.require_cuda(x, "ConvolutionModule")
        grad = A.wants_grad(self, x)
        if not grad:
            H.check_grad_mode(self)
        B, T, D = x.shape
        dev = x.device
        xc = x.contiguous()
        m8 = None
        if mask is not None:
            m2 = mask.reshape(B, T)
            valid = (m2 == 0) if self.masked_false_or_true else (m2 != 0)  # library wants 1 = keep
            m8 = valid.to(device=dev, dtype=torch.uint8).contiguous()
        if self._wv.stale(self.params(), dev):
            cw = L.ConvModWeights()
            self.fill(cw, self._wv, dev)
            self._wv.struct = cw
        if grad:
            return A.ConvModuleFunction.apply(self._wv.struct, self._act_code, A.new_dropout(self, self.after_conv[3].p), xc, m8,
                                              _chunk_size(dynchunktrain_config), *self.grad_params())
        y = torch.empty_like(xc)
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            nbytes = lib.smx_conv_module_workspace_bytes(self._wv.struct, dt, B, T)
            ws = H.workspace(dev, nbytes)
            L.check(lib.smx_conv_module_fwd(self._wv.struct, self._act_code, dt, B, T, _chunk_size(dynchunktrain_config),
                                            xc.data_ptr(), H.p_or_none(m8), None, y.data_ptr(), ws.data_ptr(),
                                            ws.numel(), H.stream_ptr(dev)))
        return y


class ConformerEncoderLayer(nn.Module):
    """x += ½·FFN1(x); x = SummaryMixing(LN1(x)) + x; x += Conv(x)·mask; x = LN2(x + ½·FFN2(x))
    (Conformer.py:518-547).  Arguments as the reference (:346-383)."""

    def __init__(
        self,
        d_model,
        d_ffn,
        nhead,
        kernel_size=31,
        kdim=None,
        vdim=None,
        activation=Swish,
        bias=True,
        dropout=0.0,
        causal=False,
        attention_type="RelPosMHAXL",
        local_proj_hid_dim=[512],
        local_proj_out_dim=512,
        summary_hid_dim=[1024],
        mode="SummaryMixing",
        use_layernorm=True,
    ):
        super().__init__()
        if attention_type != "SummaryMixing":
            raise NotImplementedError(
                f"attention_type={attention_type!r}: summarymixing_b200 builds the SummaryMixing encoder path only"
            )
        self.attention_type = attention_type
        self.mode = mode
        self.d_model = d_model
        self.mha_layer = SummaryMixing(
            enc_dim=d_model,
            nhead=nhead,
            local_proj_hid_dim=local_proj_hid_dim,
            local_proj_out_dim=local_proj_out_dim,
            summary_hid_dim=summary_hid_dim,
            summary_out_dim=d_model,
            activation=activation,
            global_dropout=dropout,
            use_layernorm=use_layernorm,
            mode=mode,
        )
        self.masked_false_or_true = False
        self.convolution_module = ConvolutionModule(
            d_model, kernel_size, bias, activation, dropout, causal=causal, masked_false_or_true=False,
        )
        self.ffn_module1 = nn.Sequential(
            nn.LayerNorm(d_model),
            PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model, dropout=dropout, activation=activation),
            nn.Dropout(dropout),
        )
        self.ffn_module2 = nn.Sequential(
            nn.LayerNorm(d_model),
            PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model, dropout=dropout, activation=activation),
            nn.Dropout(dropout),
        )
        self.norm1 = LayerNorm(d_model)
        self.norm2 = LayerNorm(d_model)
        self.drop = nn.Dropout(dropout)
        self._act_code = H.act_code(self.ffn_module1[1].ffn[1])
        self._wv = H.WeightView()

    def params(self):
        return list(self.parameters())

    @staticmethod
    def _fill_ffn(fw: L.FFNWeights, seq: nn.Sequential, wv, device):
        ln, pw = seq[0], seq[1]
        fw.ln_w = wv.ptr(ln.weight, device)
        fw.ln_b = wv.ptr(ln.bias, device)
        l1, l2 = pw.ffn[0], pw.ffn[3]
        H.fill_linear(fw.w1, wv, device, l1.weight, l1.bias, l1.in_features, l1.out_features)
        H.fill_linear(fw.w2, wv, device, l2.weight, l2.bias, l2.in_features, l2.out_features)
        H.pack_tc(wv, device, fw, L.lib().smx_ffn_packed_bytes, L.lib().smx_ffn_pack)

    def fill(self, lw: L.ConformerLayerWeights, wv: H.WeightView, device) -> None:
        self._fill_ffn(lw.ffn1, self.ffn_module1, wv, device)
        self._fill_ffn(lw.ffn2, self.ffn_module2, wv, device)
        lw.norm1_w = wv.ptr(self.norm1.norm.weight, device)
        lw.norm1_b = wv.ptr(self.norm1.norm.bias, device)
        lw.norm2_w = wv.ptr(self.norm2.norm.weight, device)
        lw.norm2_b = wv.ptr(self.norm2.norm.bias, device)
        self.mha_layer.fill(lw.cell, wv, device)
        # norm1 folded into the cell's packed image (the one-kernel cell feeds the raw rows to the tensor cores); a no-op elsewhere
        with torch.cuda.device(device):
            L.check(L.lib().smx_cell_pack_prenorm(C.byref(lw.cell), lw.norm1_w, lw.norm1_b, H.stream_ptr(device)))
        self.convolution_module.fill(lw.conv, wv, device)
        lw.act = self._act_code

    def forward(
        self,
        x,
        src_mask: Optional[torch.Tensor] = None,
        src_key_padding_mask: Optional[torch.Tensor] = None,
        pos_embs: torch.Tensor = None,
        dynchunktrain_config=None,
    ):
        """Returns (x, None) like the reference with SummaryMixing (Conformer.py:527,548)."""
        H.require_cuda(x, "ConformerEncoderLayer")
        B, T, D = x.shape
        dev = x.device
        xc = x.contiguous()
        mask = H.mask_u8(src_key_padding_mask, B, T, dev)
        smask = H.sum_mask_f32(src_mask, T, dev)
        if self._wv.stale(self.params(), dev):
            lw = L.ConformerLayerWeights()
            self.fill(lw, self._wv, dev)
            self._wv.struct = lw
        if A.wants_grad(self, x):
            return self._forward_autograd(xc, mask, smask, dynchunktrain_config), None
        H.check_grad_mode(self)
        y = torch.empty_like(xc)
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            nbytes = lib.smx_conformer_layer_workspace_bytes(self._wv.struct, dt, B, T, int(smask is not None))
            ws = H.workspace(dev, nbytes)
            L.check(lib.smx_conformer_layer_fwd(self._wv.struct, dt, B, T, _chunk_size(dynchunktrain_config),
                                                xc.data_ptr(), H.p_or_none(mask), H.p_or_none(smask), y.data_ptr(),
                                                ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        return y, None


def _ffn_params(seq: nn.Sequential):
    ln, pw = seq[0], seq[1]
    return [ln.weight, ln.bias, pw.ffn[0].weight, pw.ffn[0].bias, pw.ffn[3].weight, pw.ffn[3].bias]


def _layer_forward_autograd(self, x, mask, smask=None, dynchunktrain_config=None):
    """The layer as a chain of autograd nodes, one per libsmx module call (Conformer.py:518-547):
    FFN half-step -> norm1 -> cell + skip -> conv module + skip -> FFN half-step + norm2.  In training mode every node
    applies its dropout sites (FFN: inside PositionalwiseFeedForward and after it; cell: on the concatenation; conv module: its
    last stage); ``self.drop`` is not on the SummaryMixing path of the reference's forward."""
    lw = self._wv.struct
    d1 = A.new_dropout(self, A.same_p(self.ffn_module1[1].ffn[2].p, self.ffn_module1[2].p))
    d2 = A.new_dropout(self, A.same_p(self.ffn_module2[1].ffn[2].p, self.ffn_module2[2].p))
    x1 = A.FFNFunction.apply(lw.ffn1, self._act_code, None, d1, x, *_ffn_params(self.ffn_module1))
    n1 = A.LayerNormFunction.apply(x1, self.norm1.norm.weight, self.norm1.norm.bias, self.norm1.eps)
    x2 = self.mha_layer(n1, sum_mask=smask, src_padding_mask=mask) + x1            # (Dynamic Chunk Training: src_mask, :522-527)
    x3 = x2 + self.convolution_module(x2, mask, dynchunktrain_config=dynchunktrain_config)   # (:543-545)
    out_norm = (lw.norm2_w, lw.norm2_b, float(self.norm2.eps))
    return A.FFNFunction.apply(lw.ffn2, self._act_code, out_norm, d2, x3, *_ffn_params(self.ffn_module2),
                               self.norm2.norm.weight, self.norm2.norm.bias)


ConformerEncoderLayer._forward_autograd = _layer_forward_autograd


class ConformerEncoder(nn.Module):
    """num_layers ConformerEncoderLayers + final LayerNorm(eps=1e-6) (Conformer.py:736-763, 797-827).
    Arguments as the reference (:655-698)."""

    def __init__(
        self,
        num_layers,
        d_model,
        d_ffn,
        nhead,
        kernel_size=31,
        kdim=None,
        vdim=None,
        activation=Swish,
        bias=True,
        dropout=0.0,
        causal=False,
        attention_type="RelPosMHAXL",
        local_proj_hid_dim=[512],
        local_proj_out_dim=512,
        summary_hid_dim=[1024],
        mode="SummaryMixing",
        use_layernorm: Optional[bool] = True,
        layerdrop_prob=0.0,
        output_hidden_states=False,
    ):
        super().__init__()
        self.layers = torch.nn.ModuleList(
            [
                ConformerEncoderLayer(
                    d_ffn=d_ffn,
                    nhead=nhead,
                    d_model=d_model,
                    kdim=kdim,
                    vdim=vdim,
                    dropout=dropout,
                    activation=activation,
                    kernel_size=kernel_size,
                    bias=bias,
                    causal=causal,
                    attention_type=attention_type,
                    local_proj_hid_dim=local_proj_hid_dim,
                    local_proj_out_dim=local_proj_out_dim,
                    summary_hid_dim=summary_hid_dim,
                    use_layernorm=use_layernorm,
                    mode=mode,
                )
                for i in range(num_layers)
            ]
        )
        self.norm = LayerNorm(d_model, eps=1e-6)
        self.attention_type = attention_type
        self.layerdrop_prob = layerdrop_prob
        self.rng = np.random.default_rng()
        self.output_hidden_states = output_hidden_states
        self.d_model = d_model
        self._wv = H.WeightView()

    def params(self):
        return list(self.parameters())

    def forward(
        self,
        src,
        src_mask: Optional[torch.Tensor] = None,
        src_key_padding_mask: Optional[torch.Tensor] = None,
        pos_embs: Optional[torch.Tensor] = None,
        dynchunktrain_config=None,
    ):
        """src: (B,T,d_model).  Returns (output, attention_lst) with attention_lst = [None]*num_layers
        (or (output, hidden_lst, attention_lst) when output_hidden_states), as Conformer.py:821-827."""
        H.require_cuda(src, "ConformerEncoder")
        if A.wants_grad(self, src):
            # training path: the layers and the final norm as autograd nodes (Conformer.py:797-821, layerdrop off)
            # layerdrop (Conformer.py:798-810): one uniform draw per layer from the module's numpy generator; a layer
            # runs unless training and its draw is <= layerdrop_prob; skipped layers leave no entry in either list
            drop = self.training and self.layerdrop_prob > 0.0
            keep_probs = self.rng.random(len(self.layers)) if self.layerdrop_prob > 0.0 else None
            out, hidden_lst, attention_lst = src, [], []
            for i, layer in enumerate(self.layers):
                if drop and not keep_probs[i] > self.layerdrop_prob:
                    continue
                out, attn = layer(out, src_mask=src_mask, src_key_padding_mask=src_key_padding_mask,
                                  dynchunktrain_config=dynchunktrain_config)
                hidden_lst.append(out)
                attention_lst.append(attn)
            out = A.LayerNormFunction.apply(out, self.norm.norm.weight, self.norm.norm.bias, self.norm.eps)
            if self.output_hidden_states:
                if hidden_lst:
                    hidden_lst[-1] = out  # the last entry is the normalised output (Conformer.py:823-825)
                return out, hidden_lst, attention_lst
            return out, attention_lst
        H.check_grad_mode(self)  # layerdrop only acts in training (Conformer.py:806-810)
        B, T, D = src.shape
        dev = src.device
        xc = src.contiguous()
        mask = H.mask_u8(src_key_padding_mask, B, T, dev)
        smask = H.sum_mask_f32(src_mask, T, dev)
        n = len(self.layers)
        if self._wv.stale(self.params(), dev):
            arr = (L.ConformerLayerWeights * n)()
            for i, layer in enumerate(self.layers):
                layer.fill(arr[i], self._wv, dev)
            self._wv.struct = arr
            self._wv.norm = (self._wv.ptr(self.norm.norm.weight, dev), self._wv.ptr(self.norm.norm.bias, dev))
        y = torch.empty_like(xc)
        hidden_ptrs, hidden_lst = None, None
        if self.output_hidden_states:
            hidden_lst = [torch.empty_like(xc) for _ in range(n)]
            hidden_ptrs = (C.c_void_p * n)(*[h.data_ptr() for h in hidden_lst])
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            nbytes = lib.smx_conformer_encoder_workspace_bytes(self._wv.struct, n, dt, B, T, int(smask is not None))
            ws = H.workspace(dev, nbytes)
            L.check(lib.smx_conformer_encoder_fwd(self._wv.struct, n, self._wv.norm[0], self._wv.norm[1], dt, B, T,
                                                  _chunk_size(dynchunktrain_config), xc.data_ptr(), H.p_or_none(mask),
                                                  H.p_or_none(smask), y.data_ptr(), hidden_ptrs, ws.data_ptr(),
                                                  ws.numel(), H.stream_ptr(dev)))
        attention_lst = [None] * n
        if self.output_hidden_states:
            return y, hidden_lst, attention_lst
        return y, attention_lst
