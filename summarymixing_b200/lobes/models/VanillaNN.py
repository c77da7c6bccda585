"""ParallelLinear and VanillaNN with the reference's surface (reference: speechbrain/lobes/models/VanillaNN.py).

Same constructor signatures, error messages, parameter names/shapes and initialisation as
VanillaNN.py:58-97 and :153-196; the arithmetic runs in libsmx (smx_vanilla_nn_fwd).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
from torch import nn

from ... import _host as H
from ... import _lib as L
from ...nnet.containers import Linear


class ParallelLinear(torch.nn.Module):
    """y = x W + b applied independently on n_split slices of the feature dim (block-diagonal linear).

    weights: (n_split, input_size/n_split, n_neurons/n_split); biases: (n_split, n_neurons/n_split)
    — VanillaNN.py:85-88.  forward accepts (B,T,F) or (B,T,n_split,F/n_split) (VanillaNN.py:99-117).
    """

    def __init__(
        self,
        n_neurons,
        input_shape: Optional[list] = None,
        input_size: Optional[int] = None,
        n_split: Optional[int] = 1,
        bias: Optional[bool] = True,
        combine_out_dims: Optional[bool] = True,
    ):
        super().__init__()
        if input_size is None:
            if input_shape is None:
                raise ValueError("Expected one of input_shape or input_size")
            # a 4-D (B,T,heads,F) shape means the incoming head axis is folded into the feature dim
            input_size = math.prod(input_shape[2:]) if len(input_shape) == 4 else input_shape[-1]
        if input_size % n_split or n_neurons % n_split:
            raise ValueError("input_size and n_neurons must be dividible by n_split!")  # (sic, VanillaNN.py:80)
        self.n_split, self.combine_out_dims = n_split, combine_out_dims
        self.split_inp_dim, self.split_out_dim = input_size // n_split, n_neurons // n_split
        # head m maps input columns [m*in/h, (m+1)*in/h) to output columns [m*out/h, (m+1)*out/h)
        self.weights = nn.Parameter(torch.empty(n_split, self.split_inp_dim, self.split_out_dim))
        self.biases = nn.Parameter(torch.zeros(n_split, self.split_out_dim))
        self._reset_parameters()
        self._wv = H.WeightView()

    def _reset_parameters(self):
        # same initialisation as the reference (VanillaNN.py:92-97): kaiming-uniform on weights AND biases
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        nn.init.kaiming_uniform_(self.biases, a=math.sqrt(5))

    def _fill(self, dst: L.Linear, wv: H.WeightView, device):
        H.fill_linear(dst, wv, device, self.weights, self.biases, self.n_split * self.split_inp_dim,
                      self.n_split * self.split_out_dim, self.n_split)

    def forward(self, x):
        H.require_cuda(x, "ParallelLinear")
        B, T = x.shape[0], x.shape[1]
        xc = x.reshape(B, T, -1).contiguous()
        if xc.shape[-1] != self.n_split * self.split_inp_dim:
            raise RuntimeError(f"ParallelLinear expected {self.n_split * self.split_inp_dim} features, got {xc.shape[-1]}")
        dev = xc.device
        if self._wv.stale((self.weights, self.biases), dev):
            blk = (L.Linear * 1)()
            self._fill(blk[0], self._wv, dev)
            self._wv.struct = blk
        from ... import _autograd as A

        if A.wants_grad(self, x):  # one block, no activation, through the VanillaNN node (smx_vanilla_nn_bwd)
            y = A.VanillaNNFunction.apply(self._wv.struct, 1, L.ACT_IDENTITY, self.n_split * self.split_out_dim, xc,
                                          self.weights, self.biases)
        else:
            y = torch.empty(B, T, self.n_split * self.split_out_dim, dtype=xc.dtype, device=dev)
            _run_vanilla(self._wv.struct, 1, L.ACT_IDENTITY, xc, y)
        if not self.combine_out_dims:
            y = y.view(B, T, self.n_split, self.split_out_dim)
        return y


def _run_vanilla(blocks, n, act, xc, y):
    lib = L.lib()
    dt = H.dtype_code(xc)
    rows = xc.shape[0] * xc.shape[1]
    with torch.cuda.device(xc.device):
        nbytes = lib.smx_vanilla_nn_workspace_bytes(blocks, n, dt, rows)
        ws = H.workspace(xc.device, nbytes)
        L.check(lib.smx_vanilla_nn_fwd(blocks, n, act, dt, rows, xc.data_ptr(), y.data_ptr(), ws.data_ptr(), ws.numel(),
                                       H.stream_ptr(xc.device)))


class VanillaNN(nn.ModuleDict):
    """dnn_blocks x (linear, activation); the activation follows EVERY block, including the last
    (VanillaNN.py:168-196).  Children are named linear, act, linear_0, act_0, ... exactly as the
    SpeechBrain Sequential container names them, so state_dict keys match the reference
    (``linear.w.weight`` for dense blocks, ``linear.weights`` for split blocks)."""

    def __init__(
        self,
        input_shape,
        activation: Optional[nn.Module] = torch.nn.LeakyReLU,
        dnn_blocks: Optional[int] = 2,
        dnn_neurons: Optional[int] = 512,
        n_split: Optional[int] = 1,
    ):
        super().__init__()
        if isinstance(dnn_neurons, list):
            if len(dnn_neurons) != dnn_blocks:
                msg = "The length of the dnn_neurons list must match dnn_blocks..."
                raise ValueError(msg)
        if dnn_blocks > L.SMX_MAX_BLOCKS:
            raise NotImplementedError(f"libsmx handles at most {L.SMX_MAX_BLOCKS} blocks per VanillaNN")

        in_size = input_shape[-1]
        if len(input_shape) == 4:
            in_size = input_shape[-1] * input_shape[-2]
        self.input_size = in_size
        self.n_split = n_split
        self._linears = []
        for block_index in range(dnn_blocks):
            neurons = dnn_neurons[block_index] if isinstance(dnn_neurons, list) else dnn_neurons
            if n_split > 1:
                layer = ParallelLinear(neurons, input_size=in_size, n_split=n_split, bias=True,
                                       combine_out_dims=(block_index == dnn_blocks - 1))
            else:
                layer = Linear(neurons, input_size=in_size, bias=True)
            self._append(layer, "linear")
            self._linears.append(layer)
            self._append(activation(), "act")
            in_size = neurons
        self.output_size = in_size
        self._act_code = H.act_code(self["act"])
        self._wv = H.WeightView()

    def _append(self, layer, layer_name):
        if layer_name in self:
            index = 0
            while f"{layer_name}_{index}" in self:
                index += 1
            layer_name = f"{layer_name}_{index}"
        self.add_module(layer_name, layer)

    # -- library plumbing ---------------------------------------------------------------------
    def params(self):
        out = []
        for lin in self._linears:
            out += [lin.weights, lin.biases] if isinstance(lin, ParallelLinear) else [lin.w.weight, lin.w.bias]
        return out

    def fill(self, dst, wv: H.WeightView, device) -> int:
        """Fill an array of smx_linear with this network's blocks; returns the block count."""
        for i, lin in enumerate(self._linears):
            if isinstance(lin, ParallelLinear):
                lin._fill(dst[i], wv, device)
            else:
                H.fill_linear(dst[i], wv, device, lin.w.weight, lin.w.bias, lin.w.in_features, lin.w.out_features, 1)
        return len(self._linears)

    def forward(self, x):
        H.require_cuda(x, "VanillaNN")
        B, T = x.shape[0], x.shape[1]
        xc = x.reshape(B, T, -1).contiguous()
        if xc.shape[-1] != self.input_size:
            raise RuntimeError(f"VanillaNN expected {self.input_size} features, got {xc.shape[-1]}")
        dev = xc.device
        if self._wv.stale(self.params(), dev):
            blk = (L.Linear * L.SMX_MAX_BLOCKS)()
            self.fill(blk, self._wv, dev)
            self._wv.struct = blk
        from ... import _autograd as A

        if A.wants_grad(self, x):
            return A.VanillaNNFunction.apply(self._wv.struct, len(self._linears), self._act_code, self.output_size, xc, *self.params())
        y = torch.empty(B, T, self.output_size, dtype=xc.dtype, device=dev)
        _run_vanilla(self._wv.struct, len(self._linears), self._act_code, xc, y)
        return y
