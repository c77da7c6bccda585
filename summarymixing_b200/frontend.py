"""The acoustic frontend the recipes put in front of the encoder, with SpeechBrain's class names and constructor arguments
(conformer_summarymixing.yaml: Fbank :326-330, InputNormalization :198-200, SpectrogramDrop :298-312, Warping :315,
ConvolutionFrontEnd :145-152) and the input projection + positional encoding of TransformerASR.py:353-358, 405-406 /
Transformer.py:288-339.  The arithmetic runs in libsmx (csrc/smx_frontend.cu); CUDA only.

SpeechBrain v1.0 is not vendored with the reference: these classes restate its published behaviour ("parity unpinned",
oracle/frontend_oracle.py), including the state_dict key nesting of ConvolutionFrontEnd (convblock_i.convs.conv_0.conv.*,
convblock_i.convs.norm_0.norm.*).
"""
from __future__ import annotations

import ctypes as C
import random

import torch
import torch.nn as nn

from . import _host as H
from . import _lib as L


def _f32c(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous() if x.dtype == torch.float32 else x.float().contiguous()


class Fbank(nn.Module):
    """speechbrain.lobes.features.Fbank (deltas=False, context=False): wav (B, n_samples) -> (B, T', n_mels) log-mel dB."""

    def __init__(self, deltas=False, context=False, requires_grad=False, sample_rate=16000, f_min=0, f_max=None, n_fft=400, n_mels=40,
                 filter_shape="triangular", param_change_factor=1.0, param_rand_factor=0.0, left_frames=5, right_frames=5, win_length=25,
                 hop_length=10):
        super().__init__()
        if deltas or context or requires_grad or filter_shape != "triangular":
            raise NotImplementedError("libsmx Fbank: deltas / context / learnable or non-triangular filters are not implemented")
        self.desc = L.FbankDesc(int(sample_rate), int(n_fft), int(n_mels), float(win_length), float(hop_length), float(f_min),
                                float(f_max if f_max is not None else sample_rate / 2), 1e-10, 80.0)

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        H.require_cuda(wav, "Fbank")
        w = _f32c(wav)
        B, n = w.shape
        lib = L.lib()
        Tp = lib.smx_fbank_frames(C.byref(self.desc), n)
        out = torch.empty(B, Tp, self.desc.n_mels, dtype=torch.float32, device=w.device)
        with torch.cuda.device(w.device):
            ws = H.workspace(w.device, lib.smx_fbank_workspace_bytes(C.byref(self.desc), B))
            L.check(lib.smx_fbank_fwd(C.byref(self.desc), B, n, w.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), H.stream_ptr(w.device)))
        return out


class InputNormalization(nn.Module):
    """speechbrain.processing.features.InputNormalization(norm_type="global") at inference: (x - glob_mean) / glob_std.
    The running statistics are buffers (loaded from the recipe's `normalizer` checkpoint); updating them is the trainer's job."""

    def __init__(self, mean_norm=True, std_norm=True, norm_type="global", n_features=80):
        super().__init__()
        if norm_type != "global":
            raise NotImplementedError("libsmx InputNormalization: norm_type='global' only")
        self.register_buffer("glob_mean", torch.zeros(n_features))
        self.register_buffer("glob_std", torch.ones(n_features))

    def forward(self, x: torch.Tensor, lengths=None) -> torch.Tensor:
        H.require_cuda(x, "InputNormalization")
        xc = _f32c(x)
        y = torch.empty_like(xc)
        F = xc.shape[-1]
        with torch.cuda.device(xc.device):
            L.check(L.lib().smx_input_norm_fwd(xc.numel() // F, F, xc.data_ptr(), _f32c(self.glob_mean).data_ptr(),
                                               _f32c(self.glob_std).data_ptr(), y.data_ptr(), H.stream_ptr(xc.device)))
        return y


class SpectrogramDrop(nn.Module):
    """speechbrain.augment.freq_domain.SpectrogramDrop: `drop_count` spans of random length along time (dim=1) or frequency
    (dim=2), replaced by zeros or the batch mean.  The random draws use torch's generator exactly as upstream does."""

    def __init__(self, drop_length_low=5, drop_length_high=15, drop_count_low=1, drop_count_high=3, replace="zeros", dim=1):
        super().__init__()
        if replace not in ("zeros", "mean"):
            raise NotImplementedError("libsmx SpectrogramDrop: replace must be 'zeros' or 'mean'")
        self.drop_length_low, self.drop_length_high = drop_length_low, drop_length_high
        self.drop_count_low, self.drop_count_high = drop_count_low, drop_count_high
        self.replace, self.dim = replace, dim

    def forward(self, spectrogram: torch.Tensor) -> torch.Tensor:
        H.require_cuda(spectrogram, "SpectrogramDrop")
        x = _f32c(spectrogram).clone()
        B, T, F = x.shape
        D = x.shape[self.dim]
        n_masks = int(torch.randint(self.drop_count_low, self.drop_count_high + 1, (1,)))
        mask_len = torch.randint(self.drop_length_low, self.drop_length_high, (B, n_masks))
        mask_pos = torch.randint(0, max(1, D - int(mask_len.max())), (B, n_masks))
        self.apply_masks(x, mask_pos, mask_len)
        return x

    def apply_masks(self, x: torch.Tensor, mask_pos: torch.Tensor, mask_len: torch.Tensor) -> None:
        """In place on x (B,T,F) fp32 CUDA: the deterministic part (positions given)."""
        B, T, F = x.shape
        dev = x.device
        pos = mask_pos.to(device=dev, dtype=torch.int32).contiguous()
        ln = mask_len.to(device=dev, dtype=torch.int32).contiguous()
        lib = L.lib()
        with torch.cuda.device(dev):
            ws = H.workspace(dev, lib.smx_spec_drop_workspace_bytes())
            L.check(lib.smx_spec_drop_fwd(B, T, F, x.data_ptr(), self.dim, pos.shape[1], pos.data_ptr(), ln.data_ptr(),
                                          int(self.replace == "mean"), ws.data_ptr(), ws.numel(), H.stream_ptr(dev)))


class Warping(nn.Module):
    """speechbrain.augment.freq_domain.Warping (dim=1, bicubic): a random centre is moved by up to warp_window frames."""

    def __init__(self, warp_window=5, warp_mode="bicubic", dim=1):
        super().__init__()
        if warp_mode != "bicubic" or dim != 1:
            raise NotImplementedError("libsmx Warping: bicubic time warping (dim=1) only")
        self.warp_window = warp_window

    def forward(self, spectrogram: torch.Tensor) -> torch.Tensor:
        H.require_cuda(spectrogram, "Warping")
        T = spectrogram.shape[1]
        win = self.warp_window
        if T - win <= win:
            return spectrogram
        c = random.randrange(win, T - win)
        w = random.randrange(c - win, c + win) + 1
        return self.warp(spectrogram, c, w)

    @staticmethod
    def warp(spectrogram: torch.Tensor, c: int, w: int) -> torch.Tensor:
        x = _f32c(spectrogram)
        B, T, F = x.shape
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(L.lib().smx_time_warp_fwd(B, T, F, x.data_ptr(), int(c), int(w), y.data_ptr(), H.stream_ptr(x.device)))
        return y


class _Holder(nn.Module):
    def __init__(self, name, module):
        super().__init__()
        self.add_module(name, module)


class _ConvBlock(nn.Module):
    """Parameter holder with SpeechBrain's nesting: convs.conv_0.conv (Conv2d), convs.norm_0.norm (LayerNorm over (F', C))."""

    def __init__(self, cin, cout, kernel, fo):
        super().__init__()
        self.convs = nn.Module()
        self.convs.add_module("conv_0", _Holder("conv", nn.Conv2d(cin, cout, kernel, bias=True)))
        self.convs.add_module("norm_0", _Holder("norm", nn.LayerNorm([fo, cout])))


class ConvolutionFrontEnd(nn.Module):
    """speechbrain.lobes.models.convolution.ConvolutionFrontEnd with num_layers_per_block=1 and no residuals (the recipes'
    setting): per block Conv2d(k, stride, reflect 'same') -> LayerNorm -> LeakyReLU.  (B, T, F) -> (B, ceil(T/prod s), F'' * C)."""

    def __init__(self, input_shape, num_blocks=3, num_layers_per_block=5, out_channels=(128, 256, 512), kernel_sizes=(3, 3, 3),
                 strides=(1, 2, 2), dilations=(1, 1, 1), residuals=(True, True, True), dropout=0.15):
        super().__init__()
        if num_layers_per_block != 1 or any(residuals[:num_blocks]) or any(d != 1 for d in dilations[:num_blocks]):
            raise NotImplementedError("libsmx ConvolutionFrontEnd: num_layers_per_block=1, residuals off, dilation 1 (the recipes' configuration)")
        F, cin = int(input_shape[-1]), 1
        self.cfg = []
        for i in range(num_blocks):
            k, s, cout = int(kernel_sizes[i]), int(strides[i]), int(out_channels[i])
            fo = (F + s - 1) // s
            self.add_module(f"convblock_{i}", _ConvBlock(cin, cout, k, fo))
            self.cfg.append((F, cin, cout, k, s))
            F, cin = fo, cout
        self.out_features = F * cin

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        H.require_cuda(x, "ConvolutionFrontEnd")
        h = _f32c(x)
        B, T = h.shape[0], h.shape[1]
        lib = L.lib()
        for i, (F, cin, cout, k, s) in enumerate(self.cfg):
            blk = getattr(self, f"convblock_{i}").convs
            To, Fo = (T + s - 1) // s, (F + s - 1) // s
            y = torch.empty(B, To, Fo, cout, dtype=torch.float32, device=h.device)
            cw, cb = _f32c(blk.conv_0.conv.weight.detach()), _f32c(blk.conv_0.conv.bias.detach())
            lw, lb = _f32c(blk.norm_0.norm.weight.detach()), _f32c(blk.norm_0.norm.bias.detach())
            with torch.cuda.device(h.device):
                L.check(lib.smx_conv_frontend_block_fwd(B, T, F, cin, cout, k, s, h.data_ptr(), cw.data_ptr(), cb.data_ptr(), lw.data_ptr(),
                                                        lb.data_ptr(), y.data_ptr(), H.stream_ptr(h.device)))
            h, T = y, To
        return h.reshape(B, T, -1)


class InputProjection(nn.Module):
    """TransformerASR's custom_src_module (Linear input_size -> d_model, TransformerASR.py:353-358) followed by
    `src + positional_encoding(src)` (:405-406, Transformer.py:288-339).  The Linear keeps SpeechBrain's key (`w.weight`)."""

    def __init__(self, input_size, d_model, max_length=2500):
        super().__init__()
        self.w = nn.Linear(input_size, d_model)
        self.max_length = max_length

    def forward(self, x: torch.Tensor, out_dtype=torch.bfloat16) -> torch.Tensor:
        H.require_cuda(x, "InputProjection")
        xc = _f32c(x)
        B, T, _ = xc.shape
        D = self.w.out_features
        if T > self.max_length:
            raise RuntimeError(f"sequence length {T} exceeds the positional encoding table ({self.max_length}), as in the reference (Transformer.py:339)")
        lin = L.Linear()
        wv = H.WeightView()
        H.fill_linear(lin, wv, xc.device, self.w.weight, self.w.bias, self.w.in_features, D)
        y = torch.empty(B, T, D, dtype=out_dtype, device=xc.device)
        lib = L.lib()
        with torch.cuda.device(xc.device):
            ws = H.workspace(xc.device, lib.smx_input_proj_workspace_bytes(B, T, D))
            L.check(lib.smx_input_proj_fwd(C.byref(lin), B, T, self.max_length, xc.data_ptr(), H.dtype_code(y), y.data_ptr(), ws.data_ptr(),
                                           ws.numel(), H.stream_ptr(xc.device)))
        return y
