"""Parameter holders reproducing the SpeechBrain wrappers' state_dict key names.

The reference's modules sit on un-vendored SpeechBrain classes whose only job on this path is naming:
``Linear`` holds ``.w`` (nn.Linear), ``LayerNorm`` holds ``.norm`` (nn.LayerNorm), and
``PositionalwiseFeedForward`` holds ``.ffn`` = Sequential(Linear, act, Dropout, Linear).  These holders keep
checkpoints interchangeable; their arithmetic runs inside libsmx when the owning block's forward is called.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _host as H
from .. import _lib as L


class Linear(nn.Module):
    """speechbrain.nnet.linear.Linear: holds ``w = nn.Linear(input_size, n_neurons)``."""

    def __init__(self, n_neurons, input_shape=None, input_size=None, bias=True):
        super().__init__()
        if input_shape is None and input_size is None:
            raise ValueError("Expected one of input_shape or input_size")
        if input_size is None:
            input_size = input_shape[-1]
        self.w = nn.Linear(input_size, n_neurons, bias=bias)


class LayerNorm(nn.Module):
    """speechbrain.nnet.normalization.LayerNorm: holds ``norm = nn.LayerNorm(input_size, eps)``."""

    def __init__(self, input_size=None, input_shape=None, eps=1e-05, elementwise_affine=True):
        super().__init__()
        if input_shape is not None:
            input_size = input_shape[2:]
        if not elementwise_affine:
            raise NotImplementedError("LayerNorm without affine parameters is not used on this path")
        self.eps = eps
        self.norm = nn.LayerNorm(input_size, eps=eps)

    def forward(self, x):
        from .. import _autograd as A

        if A.wants_grad(self, x):
            return A.LayerNormFunction.apply(x, self.norm.weight, self.norm.bias, self.eps)
        return layer_norm(x, self.norm.weight, self.norm.bias, self.eps)


class PositionalwiseFeedForward(nn.Module):
    """speechbrain.nnet.attention.PositionalwiseFeedForward: ``ffn`` = [Linear, act, Dropout, Linear]."""

    def __init__(self, d_ffn, input_shape=None, input_size=None, dropout=0.0, activation=nn.ReLU):
        super().__init__()
        if input_shape is None and input_size is None:
            raise ValueError("Expected one of input_shape or input_size")
        if input_size is None:
            input_size = input_shape[-1]
        self.ffn = nn.Sequential(
            nn.Linear(input_size, d_ffn), activation(), nn.Dropout(dropout), nn.Linear(d_ffn, input_size)
        )


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> torch.Tensor:
    """nn.LayerNorm over the last dim through smx_layernorm_fwd."""
    H.require_cuda(x, "LayerNorm")
    xc = x.contiguous()
    D = xc.shape[-1]
    rows = xc.numel() // D
    y = torch.empty_like(xc)
    wv = H.WeightView()
    wv.stale((weight, bias), xc.device)
    with torch.cuda.device(xc.device):
        L.check(L.lib().smx_layernorm_fwd(H.dtype_code(xc), rows, D, xc.data_ptr(), wv.ptr(weight, xc.device),
                                          wv.ptr(bias, xc.device), float(eps), y.data_ptr(), H.stream_ptr(xc.device)))
    return y
