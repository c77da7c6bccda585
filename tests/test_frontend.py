"""Acoustic frontend: oracle cross-checks on the CPU, CUDA kernels against the oracle on the GPU.

SpeechBrain v1.0 (Fbank, InputNormalization, SpectrogramDrop, Warping, ConvolutionFrontEnd) is not vendored with the reference:
the oracle (oracle/frontend_oracle.py) restates its published behaviour -- PARITY UNPINNED for those blocks.  What CAN be pinned is:
the STFT / power stage against torch.stft, the mel + dB stages against torchaudio (the two filterbanks differ only in the falling
slope of each triangle, checked separately), the positional encoding against the reference's own class (/root/reference, build
container only), the conv block against torch.nn.functional.
"""
import math
import os

import pytest
import torch

from oracle import frontend_oracle as FO


def _wav(B=2, n=16000, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 16000.0
    tones = sum(torch.sin(2 * math.pi * f * t) * a for f, a in ((220.0, 0.5), (1330.0, 0.2), (3700.0, 0.1)))
    return tones[None] * torch.tensor([1.0, 0.3])[:B, None] + 0.01 * torch.randn(B, n, generator=g)


def test_oracle_stft_and_mel_against_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    wav = _wav()
    p = FO.stft_power(wav)
    spec = torchaudio.transforms.Spectrogram(n_fft=512, win_length=512, hop_length=160, window_fn=torch.hamming_window, power=2.0,
                                             center=True, pad_mode="constant")(wav).transpose(1, 2)
    assert p.shape == spec.shape == (2, 101, 257)
    assert float((p - spec).abs().max()) <= 1e-3 * float(spec.abs().max())
    fb = FO.mel_filterbank()
    ta = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 80, 16000, norm=None, mel_scale="htk")
    # same centres and same rising slopes; SpeechBrain's triangles fall with the rising slope's width (symmetric in Hz)
    assert fb.shape == ta.shape == (257, 80)
    peak_fb, peak_ta = fb.argmax(0), ta.argmax(0)
    assert int((peak_fb - peak_ta).abs().max()) <= 1
    mel = torch.linspace(0.0, 2595.0 * math.log10(1.0 + 8000.0 / 700.0), 82, dtype=torch.float64)
    f_central = (700.0 * (10.0 ** (mel / 2595.0) - 1.0))[1:-1]
    freqs = torch.linspace(0, 8000, 257, dtype=torch.float64)
    rising = (freqs[:, None] < f_central[None, :]) & (ta > 0)
    assert float((fb - ta)[rising].abs().max()) < 1e-4
    db = FO.fbank(wav)
    assert db.shape == (2, 101, 80)
    assert float(db.amax(dim=(1, 2)).min() - db.amin(dim=(1, 2)).max()) <= 80.0 + 1e-4  # top_db clamp per utterance


@pytest.mark.refshim
def test_oracle_positional_encoding_matches_the_reference_class():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "sbshim"))
    from speechbrain.lobes.models.transformer.Transformer import PositionalEncoding

    x = torch.zeros(2, 333, 256)
    assert torch.equal(PositionalEncoding(256)(x), FO.positional_encoding(333, 256))


def test_oracle_conv_frontend_shapes():
    import summarymixing_b200 as S

    fe = S.frontend.ConvolutionFrontEnd((8, 10, 80), 2, 1, (64, 32), (3, 3), (2, 2), (1, 1), (False, False))
    y = FO.conv_frontend(torch.randn(2, 37, 80), dict(fe.state_dict()))
    assert y.shape == (2, 10, 640) and fe.out_features == 640


def test_oracle_spec_drop_and_warp_identities():
    x = torch.randn(2, 50, 8, generator=torch.Generator().manual_seed(1))
    pos, ln = torch.tensor([[3, 20], [0, 45]]), torch.tensor([[4, 5], [2, 5]])
    y = FO.spectrogram_drop(x, pos, ln, dim=1, replace="mean")
    assert torch.allclose(y[0, 3:7], x.mean().expand(4, 8)) and torch.equal(y[0, 7:20], x[0, 7:20])
    assert torch.allclose(FO.time_warp(x, 25, 25), x, atol=1e-6)  # centre not moved: identity


# ---------------------------------------------------------------------------------------------------- GPU
DEV = "cuda:0"


@pytest.mark.gpu
def test_fbank_kernel_vs_oracle():
    import summarymixing_b200 as S

    wav = _wav(2, 16000 * 2 + 37, seed=3)
    ref = FO.fbank(wav)
    fb = S.frontend.Fbank(sample_rate=16000, n_fft=512, n_mels=80, win_length=32)
    y = fb(wav.to(DEV)).cpu()
    assert y.shape == ref.shape
    # dB of a power spectrum: fp32 FFT vs torch.stft differ by ~1e-6 relative in power, i.e. ~1e-5 dB; bins at the amin floor excepted
    assert float((y - ref).abs().max()) < 5e-3, float((y - ref).abs().max())


@pytest.mark.gpu
def test_input_norm_drop_warp_kernels_vs_oracle():
    import summarymixing_b200 as S

    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 211, 80, generator=g) * 7 - 30
    norm = S.frontend.InputNormalization()
    norm.glob_mean.copy_(torch.randn(80, generator=g))
    norm.glob_std.copy_(torch.rand(80, generator=g) + 0.5)
    y = norm.to(DEV)(x.to(DEV)).cpu()
    assert float((y - FO.input_norm(x, norm.glob_mean.cpu(), norm.glob_std.cpu())).abs().max()) < 1e-5
    for dim, hi in ((1, 211), (2, 80)):
        drop = S.frontend.SpectrogramDrop(10, 20, 4, 4, replace="mean", dim=dim)
        pos = torch.randint(0, hi - 20, (3, 4), generator=g)
        ln = torch.randint(10, 20, (3, 4), generator=g)
        xd = x.clone().to(DEV)
        drop.apply_masks(xd, pos, ln)
        assert float((xd.cpu() - FO.spectrogram_drop(x, pos, ln, dim=dim, replace="mean")).abs().max()) < 1e-4
    for c, w in ((100, 103), (57, 53), (6, 10)):
        yw = S.frontend.Warping.warp(x.to(DEV), c, w).cpu()
        assert float((yw - FO.time_warp(x, c, w)).abs().max()) < 2e-4, (c, w)


@pytest.mark.gpu
def test_conv_frontend_and_input_projection_vs_oracle():
    import summarymixing_b200 as S
    from oracle.seeded import fill_module

    fe = S.frontend.ConvolutionFrontEnd((8, 10, 80), 2, 1, (64, 32), (3, 3), (2, 2), (1, 1), (False, False))
    fill_module(fe, 17)
    x = torch.randn(2, 203, 80, generator=torch.Generator().manual_seed(6))
    ref = FO.conv_frontend(x, dict(fe.state_dict()))
    y = fe.to(DEV)(x.to(DEV)).cpu()
    assert y.shape == ref.shape == (2, 51, 640)
    assert float((y - ref).abs().max()) < 2e-4 * max(1.0, float(ref.abs().max()))
    proj = S.frontend.InputProjection(640, 256)
    fill_module(proj, 18)
    pref = FO.input_projection(ref, proj.w.weight.detach(), proj.w.bias.detach())
    py = proj.to(DEV)(ref.to(DEV), out_dtype=torch.float32).cpu()
    assert float((py - pref).abs().max()) < 5e-4 * max(1.0, float(pref.abs().max()))
    with pytest.raises(RuntimeError):  # the reference's table has 2500 rows (Transformer.py:339)
        proj(torch.zeros(1, 2501, 640, device=DEV))


@pytest.mark.gpu
def test_waveform_to_encoder_end_to_end():
    """fbank -> normalise -> CNN frontend -> projection + positional encoding -> 2-layer encoder, against the oracle chain."""
    import summarymixing_b200 as S
    from oracle import smx_oracle as O
    from oracle.seeded import fill_module

    wav = _wav(2, 16000 * 3, seed=9)
    fb = S.frontend.Fbank(sample_rate=16000, n_fft=512, n_mels=80, win_length=32)
    fe = S.frontend.ConvolutionFrontEnd((8, 10, 80), 2, 1, (64, 32), (3, 3), (2, 2), (1, 1), (False, False))
    proj = S.frontend.InputProjection(640, 256)
    enc = S.ConformerEncoder(2, 256, 1024, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[256], local_proj_out_dim=256,
                             summary_hid_dim=[256]).eval()
    for i, m in enumerate((fe, proj, enc)):
        fill_module(m, 30 + i)
    feats = FO.fbank(wav)
    mean, std = feats.mean(dim=(0, 1)), feats.std(dim=(0, 1))
    h = FO.conv_frontend(FO.input_norm(feats, mean, std), dict(fe.state_dict()))
    src = FO.input_projection(h, proj.w.weight.detach(), proj.w.bias.detach())
    T = src.shape[1]
    mask = torch.arange(T)[None] < torch.tensor([T, T * 2 // 3])[:, None]
    ref = O.conformer_encoder(src, dict(enc.state_dict()), 2, act="swish", src_key_padding_mask=mask)
    norm = S.frontend.InputNormalization()
    norm.glob_mean.copy_(mean)
    norm.glob_std.copy_(std)
    with torch.no_grad():
        s = proj.to(DEV)(fe.to(DEV)(norm.to(DEV)(fb(wav.to(DEV)))), out_dtype=torch.float32)
        y = enc.to(DEV)(s, src_key_padding_mask=mask.to(DEV))[0].cpu()
    err = float((y - ref).abs().max())
    assert err < 2e-3 * max(1.0, float(ref.abs().max())), err
