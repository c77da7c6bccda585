"""speechbrain.lobes.models.convolution stand-in: ConvolutionalSpatialGatingUnit (CSGU).

PARITY UNPINNED: restated from upstream SpeechBrain v1.0 behaviour, which is not
available here.  Semantics: split channels in two halves; LayerNorm the gate half; depthwise conv
over time, 'same' length with reflect padding ((k-1)/2 each side), conv weight ~ N(0, 1e-6), conv
bias = 1; optional Linear (weight ~ N(0,1e-6), bias = 1); gate activation; multiply by the other
half; dropout.  Upstream nests the wrapped modules as ``norm.norm`` and ``conv.conv``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _SBConv1dSameReflect(nn.Module):
    """Depthwise speechbrain.nnet.CNN.Conv1d(padding='same', padding_mode='reflect') on (B,T,C)."""

    def __init__(self, channels, kernel_size):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv = nn.Conv1d(channels, channels, kernel_size, stride=1, padding=0, groups=channels, bias=True)

    def forward(self, x):
        x = x.transpose(1, -1)
        pad = (self.kernel_size - 1) // 2
        x = F.pad(x, (pad, pad), mode="reflect")
        x = self.conv(x)
        return x.transpose(1, -1)


class _SBLayerNorm(nn.Module):
    def __init__(self, size):
        super().__init__()
        self.norm = nn.LayerNorm(size)

    def forward(self, x):
        return self.norm(x)


class ConvolutionalSpatialGatingUnit(nn.Module):
    def __init__(self, input_size, kernel_size=31, dropout=0.0, use_linear_after_conv=False, activation=nn.Identity):
        super().__init__()
        self.input_size = input_size
        self.use_linear_after_conv = use_linear_after_conv
        self.activation = activation()
        if self.input_size % 2 != 0:
            raise ValueError("Input size must be divisible by 2!")
        n_channels = input_size // 2
        self.norm = _SBLayerNorm(n_channels)
        self.conv = _SBConv1dSameReflect(n_channels, kernel_size)
        self.linear = None
        if use_linear_after_conv:
            self.linear = nn.Linear(n_channels, n_channels)
            nn.init.normal_(self.linear.weight, std=1e-6)
            nn.init.ones_(self.linear.bias)
        nn.init.normal_(self.conv.conv.weight, std=1e-6)
        nn.init.ones_(self.conv.conv.bias)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        x1, x2 = x.chunk(2, dim=-1)
        x2 = self.norm(x2)
        x2 = self.conv(x2)
        if self.linear is not None:
            x2 = self.linear(x2)
        x2 = self.activation(x2)
        return self.dropout(x2 * x1)
