"""Branchformer encoder blocks around the SummaryMixing cell, with the reference's surface
(reference: speechbrain/lobes/models/transformer/Branchformer.py).

ConvolutionBranch (:31-97), BranchformerEncoderLayer (:100-334) and BranchformerEncoder (:337-491).
The ConvolutionalSpatialGatingUnit is SpeechBrain's (un-vendored): its state_dict keys
(csgu.norm.norm.*, csgu.conv.conv.*, csgu.linear.*) and arithmetic follow upstream v1.0 — parity unpinned
beyond the oracle's restatement.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .... import _autograd as A
from .... import _host as H
from .... import _lib as L
from ....nnet.containers import LayerNorm
from ....nnet.summary_mixing import SummaryMixing
from ...models.VanillaNN import VanillaNN


class _Wrapped(nn.Module):
    """Holds one module under a fixed attribute name (SpeechBrain's wrapper nesting: norm.norm, conv.conv)."""

    def __init__(self, name, module):
        super().__init__()
        self.add_module(name, module)


class ConvolutionalSpatialGatingUnit(nn.Module):
    """Parameter holder for SpeechBrain's CSGU: LayerNorm(U/2) on the gate half, depthwise Conv1d
    ('same', reflect padding), optional Linear, gate activation; initialisation as upstream
    (conv/linear weight ~ N(0,1e-6), bias = 1)."""

    def __init__(self, input_size, kernel_size=31, dropout=0.0, use_linear_after_conv=False, activation=nn.Identity):
        super().__init__()
        if input_size % 2 != 0:
            raise ValueError("Input size must be divisible by 2!")
        n_channels = input_size // 2
        self.input_size = input_size
        self.kernel_size = kernel_size
        self.use_linear_after_conv = use_linear_after_conv
        self.activation = activation()
        self.norm = _Wrapped("norm", nn.LayerNorm(n_channels))
        self.conv = _Wrapped(
            "conv", nn.Conv1d(n_channels, n_channels, kernel_size, stride=1, padding=0, groups=n_channels, bias=True)
        )
        if use_linear_after_conv:
            self.linear = nn.Linear(n_channels, n_channels)
            nn.init.normal_(self.linear.weight, std=1e-6)
            nn.init.ones_(self.linear.bias)
        nn.init.normal_(self.conv.conv.weight, std=1e-6)
        nn.init.ones_(self.conv.conv.bias)
        self.dropout = nn.Dropout(dropout)


class ConvolutionBranch(nn.Module):
    """Channel proj -> act -> CSGU -> channel proj (Branchformer.py:86-97).  Arguments as the reference (:37-52)."""

    def __init__(
        self,
        input_size,
        linear_units=3072,
        kernel_size=31,
        activation=nn.GELU,
        gate_activation=nn.Identity,
        dropout=0.0,
        use_linear_after_conv=False,
    ):
        super().__init__()
        self.pre_channel_proj = nn.Linear(input_size, linear_units)
        self.post_channel_proj = nn.Linear(linear_units // 2, input_size)
        self.activation = activation()
        self.csgu = ConvolutionalSpatialGatingUnit(
            input_size=linear_units,
            kernel_size=kernel_size,
            dropout=dropout,
            use_linear_after_conv=use_linear_after_conv,
            activation=gate_activation,
        )
        self._act_code = H.act_code(self.activation)
        self._gate_code = H.act_code(self.csgu.activation)
        self._wv = H.WeightView()

    def grad_params(self):
        """Parameters in the order smx_convbranch_grads lists their gradients."""
        out = [self.pre_channel_proj.weight, self.pre_channel_proj.bias, self.post_channel_proj.weight, self.post_channel_proj.bias,
               self.csgu.norm.norm.weight, self.csgu.norm.norm.bias, self.csgu.conv.conv.weight, self.csgu.conv.conv.bias]
        if self.csgu.use_linear_after_conv:
            out += [self.csgu.linear.weight, self.csgu.linear.bias]
        return out

    def forward(self, x):
        """(B,T,input_size) -> (B,T,input_size), Branchformer.py:86-97.  Runs on the fp32-math arm (linears as split-bf16 tensor-core
        GEMMs); recorded by autograd (smx_conv_branch_train_bwd) when x requires grad or in training mode, where the CSGU's dropout
        applies.  Inside a BranchformerEncoderLayer's inference forward the branch is part of the layer's fused path instead."""
        H.require_cuda(x, "ConvolutionBranch")
        dev = x.device
        if self._wv.stale(list(self.parameters()), dev):
            bw = L.ConvBranchWeights()
            self.fill(bw, self._wv, dev)
            self._wv.struct = bw
        return self.run(self._wv.struct, x)

    def run(self, bw: L.ConvBranchWeights, x):
        """The branch on a filled weights struct (its own, or the one inside the owning layer's struct)."""
        drop = A.new_dropout(self, self.csgu.dropout.p)
        if A.wants_grad(self, x):
            return A.ConvBranchFunction.apply(bw, drop, x, *self.grad_params())
        with torch.no_grad():
            return A.ConvBranchFunction.apply(bw, drop, x, *self.grad_params())

    def fill(self, bw: L.ConvBranchWeights, wv: H.WeightView, device) -> None:
        pre, post = self.pre_channel_proj, self.post_channel_proj
        H.fill_linear(bw.pre, wv, device, pre.weight, pre.bias, pre.in_features, pre.out_features)
        H.fill_linear(bw.post, wv, device, post.weight, post.bias, post.in_features, post.out_features)
        bw.csgu_ln_w = wv.ptr(self.csgu.norm.norm.weight, device)
        bw.csgu_ln_b = wv.ptr(self.csgu.norm.norm.bias, device)
        bw.csgu_dw_w = wv.ptr(self.csgu.conv.conv.weight, device)
        bw.csgu_dw_b = wv.ptr(self.csgu.conv.conv.bias, device)
        if self.csgu.use_linear_after_conv:
            lin = self.csgu.linear
            H.fill_linear(bw.csgu_linear, wv, device, lin.weight, lin.bias, lin.in_features, lin.out_features)
        bw.kernel_size = self.csgu.kernel_size
        bw.act = self._act_code
        bw.gate_act = self._gate_code


class BranchformerEncoderLayer(nn.Module):
    """x + merge_proj(cat[SummaryMixing(LN(x)), ConvBranch(LN(x))]) (Branchformer.py:262-281).
    Arguments as the reference (:103-146)."""

    def __init__(
        self,
        d_model,
        nhead,
        kernel_size=31,
        kdim=None,
        vdim=None,
        activation=nn.GELU,
        dropout=0.0,
        attention_type="SummaryMixing",
        csgu_linear_units=3072,
        gate_activation=nn.Identity,
        use_linear_after_conv=False,
        local_proj_hid_dim=[512],
        local_proj_out_dim=512,
        summary_hid_dim=[1024],
        summary_out_dim=1024,
        mode="SummaryMixing",
    ):
        super().__init__()
        if attention_type != "SummaryMixing":
            raise NotImplementedError(
                f"attention_type={attention_type!r}: summarymixing_b200 builds the SummaryMixing encoder path only"
            )
        self.attention_type = attention_type
        self.mode = mode
        self.mha_layer = SummaryMixing(
            enc_dim=d_model,
            nhead=nhead,
            local_proj_hid_dim=local_proj_hid_dim,
            local_proj_out_dim=local_proj_out_dim,
            summary_hid_dim=summary_hid_dim,
            summary_out_dim=summary_out_dim,
            activation=activation,
            mode=mode,
        )
        self.merge_dnn_blocks = summary_hid_dim + [d_model]
        self.merge_proj = VanillaNN(
            input_shape=[None, None, local_proj_out_dim + summary_out_dim],
            dnn_blocks=len(self.merge_dnn_blocks),
            dnn_neurons=self.merge_dnn_blocks,
            activation=activation,
        )
        self.norm_mhsa = LayerNorm(d_model)
        self.convolution_branch = ConvolutionBranch(
            input_size=d_model,
            kernel_size=kernel_size,
            linear_units=csgu_linear_units,
            activation=activation,
            gate_activation=gate_activation,
            dropout=dropout,
            use_linear_after_conv=use_linear_after_conv,
        )
        self.norm_conv = LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self._act_code = H.act_code(self.convolution_branch.activation)
        self._wv = H.WeightView()

    def params(self):
        return list(self.parameters())

    def fill(self, lw: L.BranchformerLayerWeights, wv: H.WeightView, device) -> None:
        lw.norm_mhsa_w = wv.ptr(self.norm_mhsa.norm.weight, device)
        lw.norm_mhsa_b = wv.ptr(self.norm_mhsa.norm.bias, device)
        lw.norm_conv_w = wv.ptr(self.norm_conv.norm.weight, device)
        lw.norm_conv_b = wv.ptr(self.norm_conv.norm.bias, device)
        self.mha_layer.fill(lw.cell, wv, device)
        self.convolution_branch.fill(lw.branch, wv, device)
        lw.n_merge = self.merge_proj.fill(lw.merge, wv, device)
        lw.act = self._act_code
        H.pack_tc(wv, device, lw, L.lib().smx_branchformer_packed_bytes, L.lib().smx_branchformer_pack)

    def forward(
        self,
        x,
        src_mask: Optional[torch.Tensor] = None,
        src_key_padding_mask: Optional[torch.Tensor] = None,
        pos_embs: Optional[torch.Tensor] = None,
    ):
        H.require_cuda(x, "BranchformerEncoderLayer")
        B, T, D = x.shape
        dev = x.device
        xc = x.contiguous()
        mask = H.mask_u8(src_key_padding_mask, B, T, dev)
        smask = H.sum_mask_f32(src_mask, T, dev)
        if self._wv.stale(self.params(), dev):
            lw = L.BranchformerLayerWeights()
            self.fill(lw, self._wv, dev)
            self._wv.struct = lw
        if A.wants_grad(self, x):
            return self._forward_autograd(xc, mask, smask), None
        H.check_grad_mode(self)
        y = torch.empty_like(xc)
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            nbytes = lib.smx_branchformer_layer_workspace_bytes(self._wv.struct, dt, B, T, int(smask is not None))
            ws = H.workspace(dev, nbytes)
            L.check(lib.smx_branchformer_layer_fwd(self._wv.struct, dt, B, T, xc.data_ptr(), H.p_or_none(mask),
                                                   H.p_or_none(smask), y.data_ptr(), ws.data_ptr(), ws.numel(),
                                                   H.stream_ptr(dev)))
        return y, None


def _branchformer_layer_forward_autograd(self, x, mask, smask=None):
    """The layer as a chain of autograd nodes (Branchformer.py:262-334): norm_mhsa -> cell -> dropout | norm_conv -> convolution
    branch -> dropout | merge_proj(cat) -> dropout -> + x.  In training mode the three nn.Dropout calls of the layer use one
    smx_dropout (sites 1, 2, 3: counter-based masks, regenerated by the backward) and the CSGU's dropout another (site 0)."""
    drop = A.new_dropout(self, self.dropout.p)
    x1 = self.mha_layer(self.norm_mhsa(x), sum_mask=smask, src_padding_mask=mask)       # :317-322
    x1 = A.dropout(drop, 1, x1)                                                          # :334
    x2 = self.convolution_branch.run(self._wv.struct.branch, self.norm_conv(x))          # :292-293 (no mask, :276)
    x2 = A.dropout(drop, 2, x2)                                                          # :294
    merged = self.merge_proj(torch.cat([x1, x2], dim=-1))                                # :279, :220-226
    return x + A.dropout(drop, 3, merged)


BranchformerEncoderLayer._forward_autograd = _branchformer_layer_forward_autograd


class BranchformerEncoder(nn.Module):
    """num_layers BranchformerEncoderLayers + final LayerNorm(eps=1e-6) (Branchformer.py:421-445, 479-491).
    Arguments as the reference (:340-385)."""

    def __init__(
        self,
        num_layers,
        d_model,
        nhead,
        kernel_size=31,
        kdim=None,
        vdim=None,
        activation=nn.GELU,
        dropout=0.0,
        attention_type="SummaryMixing",
        csgu_linear_units=3072,
        gate_activation=nn.Identity,
        use_linear_after_conv=False,
        local_proj_hid_dim=[512],
        local_proj_out_dim=512,
        summary_hid_dim=[1024],
        summary_out_dim=1024,
        mode="SummaryMixing",
    ):
        super().__init__()
        self.layers = torch.nn.ModuleList(
            [
                BranchformerEncoderLayer(
                    nhead=nhead,
                    d_model=d_model,
                    kdim=kdim,
                    vdim=vdim,
                    dropout=dropout,
                    activation=activation,
                    kernel_size=kernel_size,
                    attention_type=attention_type,
                    csgu_linear_units=csgu_linear_units,
                    gate_activation=gate_activation,
                    use_linear_after_conv=use_linear_after_conv,
                    local_proj_hid_dim=local_proj_hid_dim,
                    local_proj_out_dim=local_proj_out_dim,
                    summary_hid_dim=summary_hid_dim,
                    summary_out_dim=summary_out_dim,
                    mode=mode,
                )
                for i in range(num_layers)
            ]
        )
        self.norm = LayerNorm(d_model, eps=1e-6)
        self.attention_type = attention_type
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
        assert dynchunktrain_config is None, "Dynamic Chunk Training unsupported for this encoder"
        H.require_cuda(src, "BranchformerEncoder")
        B, T, D = src.shape
        dev = src.device
        xc = src.contiguous()
        mask = H.mask_u8(src_key_padding_mask, B, T, dev)
        smask = H.sum_mask_f32(src_mask, T, dev)
        n = len(self.layers)
        if A.wants_grad(self, src):  # training / differentiable path: the layers' autograd chains, then the final norm (:479-491)
            out = xc
            for layer in self.layers:
                out, _ = layer(out, src_mask=src_mask, src_key_padding_mask=mask)
            return self.norm(out), [None] * n
        H.check_grad_mode(self)
        if self._wv.stale(self.params(), dev):
            arr = (L.BranchformerLayerWeights * n)()
            for i, layer in enumerate(self.layers):
                layer.fill(arr[i], self._wv, dev)
            self._wv.struct = arr
            self._wv.norm = (self._wv.ptr(self.norm.norm.weight, dev), self._wv.ptr(self.norm.norm.bias, dev))
        y = torch.empty_like(xc)
        lib = L.lib()
        dt = H.dtype_code(xc)
        with torch.cuda.device(dev):
            nbytes = lib.smx_branchformer_encoder_workspace_bytes(self._wv.struct, n, dt, B, T, int(smask is not None))
            ws = H.workspace(dev, nbytes)
            L.check(lib.smx_branchformer_encoder_fwd(self._wv.struct, n, self._wv.norm[0], self._wv.norm[1], dt, B, T,
                                                     xc.data_ptr(), H.p_or_none(mask), H.p_or_none(smask), y.data_ptr(),
                                                     ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))
        return y, [None] * n
