"""CPU restatement of the acoustic frontend the recipes put in front of the encoder (TEST INFRASTRUCTURE, like smx_oracle.py).

The reference repository configures these blocks only by YAML name (recipes/LibriSpeech/ASR/transformer/hparams/
conformer_summarymixing.yaml: Fbank :326-330, InputNormalization :198-200, SpectrogramDrop :298-312, Warping :315,
ConvolutionFrontEnd :145-152) and applies the input projection + positional encoding in TransformerASR.py:353-358, 405-406 /
Transformer.py:288-339.  The classes themselves live in SpeechBrain v1.0, which is NOT vendored in /root/reference and cannot be
installed here: **parity unpinned** -- every function below restates SpeechBrain's published behaviour (speechbrain/processing/
features.py, speechbrain/augment/freq_domain.py, speechbrain/lobes/models/convolution.py, speechbrain/nnet/CNN.py) and is
cross-checked against torch.stft / torchaudio where a stage has an independent implementation (tests/test_frontend.py).
Only Transformer.py's PositionalEncoding is in the reference tree and is checked against it.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---- Fbank = STFT -> power spectrum -> triangular mel filterbank -> dB with top_db clamp ---------------------------------
def stft_power(wav: Tensor, sample_rate: int = 16000, n_fft: int = 512, win_length_ms: float = 32.0, hop_length_ms: float = 10.0) -> Tensor:
    """speechbrain.processing.features.STFT (hamming window, center=True, pad_mode='constant', not normalised, one-sided)
    followed by spectral_magnitude(power=1) = re^2 + im^2.  wav (B, n) -> (B, T', n_fft/2 + 1), T' = 1 + n // hop."""
    win = int(round(sample_rate / 1000.0 * win_length_ms))
    hop = int(round(sample_rate / 1000.0 * hop_length_ms))
    window = torch.hamming_window(win, dtype=wav.dtype)
    s = torch.stft(wav, n_fft, hop, win, window, center=True, pad_mode="constant", normalized=False, onesided=True, return_complex=True)
    return (s.real ** 2 + s.imag ** 2).transpose(1, 2)


def mel_filterbank(n_mels: int = 80, n_fft: int = 512, sample_rate: int = 16000, f_min: float = 0.0, f_max: Optional[float] = None,
                   dtype=torch.float32) -> Tensor:
    """speechbrain.processing.features.Filterbank, triangular filters: centres equally spaced on the mel scale
    (mel = 2595 log10(1 + f / 700)); filter i rises and falls with the SAME slope 1 / (hz[i+1] - hz[i]) (the band to the left of
    its centre) around f_central = hz[i+1].  (n_fft/2 + 1, n_mels)."""
    f_max = sample_rate / 2 if f_max is None else f_max
    n_stft = n_fft // 2 + 1
    to_mel = lambda hz: 2595.0 * math.log10(1.0 + hz / 700.0)  # noqa: E731
    mel = torch.linspace(to_mel(f_min), to_mel(f_max), n_mels + 2, dtype=torch.float64)
    hz = 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    band = (hz[1:] - hz[:-1])[:-1]
    f_central = hz[1:-1]
    all_freqs = torch.linspace(0, sample_rate // 2, n_stft, dtype=torch.float64)
    slope = (all_freqs[None, :] - f_central[:, None]) / band[:, None]
    fb = torch.clamp(torch.minimum(slope + 1.0, -slope + 1.0), min=0.0)
    return fb.T.to(dtype)


def fbank(wav: Tensor, sample_rate: int = 16000, n_fft: int = 512, n_mels: int = 80, win_length_ms: float = 32.0,
          hop_length_ms: float = 10.0, f_min: float = 0.0, f_max: Optional[float] = None, amin: float = 1e-10, top_db: float = 80.0) -> Tensor:
    """speechbrain.lobes.features.Fbank (deltas=False, context=False): (B, n) -> (B, T', n_mels) log-mel energies in dB,
    clamped per utterance to [max - top_db, max]."""
    p = stft_power(wav, sample_rate, n_fft, win_length_ms, hop_length_ms)
    m = p @ mel_filterbank(n_mels, n_fft, sample_rate, f_min, f_max, dtype=wav.dtype)
    db = 10.0 * torch.log10(torch.clamp(m, min=amin))
    floor = db.amax(dim=(-2, -1), keepdim=True) - top_db
    return torch.maximum(db, floor)


def input_norm(x: Tensor, mean: Tensor, std: Tensor) -> Tensor:
    """speechbrain.processing.features.InputNormalization(norm_type='global') at inference: (x - glob_mean) / glob_std per feature."""
    return (x - mean) / std


# ---- SpecAugment blocks (training) ----------------------------------------------------------------------------------------
def spectrogram_drop(x: Tensor, mask_pos: Tensor, mask_len: Tensor, dim: int = 1, replace: str = "mean") -> Tensor:
    """speechbrain.augment.freq_domain.SpectrogramDrop with the random draws made by the caller: mask_pos / mask_len (B, n_masks)
    ints; positions [pos, pos + len) along `dim` (1 = time, 2 = frequency) of utterance b are replaced by 0 or by the mean of
    the WHOLE batch tensor."""
    B, T, Fd = x.shape
    D = x.shape[dim]
    ar = torch.arange(D).view(1, 1, -1)
    mask = ((mask_pos.unsqueeze(2) <= ar) & (ar < (mask_pos + mask_len).unsqueeze(2))).any(dim=1)
    mask = mask.unsqueeze(2) if dim == 1 else mask.unsqueeze(1)
    val = 0.0 if replace == "zeros" else float(x.mean())
    return x.masked_fill(mask, val)


def time_warp(x: Tensor, c: int, w: int) -> Tensor:
    """speechbrain.augment.freq_domain.Warping (dim=1, bicubic) with the random centre c and its new position w chosen by the
    caller: frames [0, c) are resampled to w frames, frames [c, T) to T - w frames (bicubic, align_corners=True)."""
    B, T, Fd = x.shape
    x4 = x.unsqueeze(1)
    left = F.interpolate(x4[:, :, :c], (w, Fd), mode="bicubic", align_corners=True)
    right = F.interpolate(x4[:, :, c:], (T - w, Fd), mode="bicubic", align_corners=True)
    return torch.cat([left, right], dim=2).squeeze(1)


# ---- ConvolutionFrontEnd ------------------------------------------------------------------------------------------------------
def conv_frontend(x: Tensor, sd: dict, prefix: str = "", num_blocks: int = 2, strides=(2, 2)) -> Tensor:
    """speechbrain.lobes.models.convolution.ConvolutionFrontEnd with num_layers_per_block=1, residuals off (the recipes'
    setting): per block Conv2d(k=3, stride s, 'same' padding = reflect-pad (k-1)/2 on both sides of time and frequency)
    -> LayerNorm over (F', C) jointly -> LeakyReLU(0.01) -> (dropout).  x (B, T, F) -> (B, T'', F'' * C)."""
    h = x.unsqueeze(-1)  # (B, T, F, 1) channels-last like SpeechBrain's Conv2d
    for i in range(num_blocks):
        w = sd[f"{prefix}convblock_{i}.convs.conv_0.conv.weight"]
        b = sd[f"{prefix}convblock_{i}.convs.conv_0.conv.bias"]
        k = w.shape[-1]
        hc = h.permute(0, 3, 1, 2)  # (B, C, T, F)
        hc = F.pad(hc, (k // 2, k // 2, k // 2, k // 2), mode="reflect")
        hc = F.conv2d(hc, w, b, stride=strides[i])
        h = hc.permute(0, 2, 3, 1)  # (B, T', F', C)
        lw = sd[f"{prefix}convblock_{i}.convs.norm_0.norm.weight"]
        lb = sd[f"{prefix}convblock_{i}.convs.norm_0.norm.bias"]
        h = F.layer_norm(h, h.shape[2:], lw, lb, 1e-5)
        h = F.leaky_relu(h, 0.01)
    return h.reshape(h.shape[0], h.shape[1], -1)


# ---- input projection + positional encoding -------------------------------------------------------------------------------------
def positional_encoding(T: int, D: int, max_len: int = 2500, dtype=torch.float32) -> Tensor:
    """Transformer.py:288-339 (PositionalEncoding): pe[t, 2i] = sin(t / 10000^(2i/D)), pe[t, 2i+1] = cos(...); (1, T, D)."""
    if T > max_len:
        raise RuntimeError(f"sequence length {T} exceeds the positional table ({max_len}), Transformer.py:339")
    pos = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
    den = torch.exp(torch.arange(0, D, 2, dtype=torch.float32) * (-(math.log(10000.0) / D)))
    pe = torch.zeros(T, D, dtype=torch.float32)
    pe[:, 0::2] = torch.sin(pos * den)
    pe[:, 1::2] = torch.cos(pos * den)
    return pe.unsqueeze(0).to(dtype)


def input_projection(x: Tensor, weight: Tensor, bias: Tensor, max_len: int = 2500) -> Tensor:
    """TransformerASR.py:353-358 (custom_src_module: Linear input_size -> d_model, dropout off) and :405-406 (src = src +
    positional_encoding(src))."""
    y = x @ weight.T + bias
    return y + positional_encoding(y.shape[1], y.shape[2], max_len, y.dtype)
