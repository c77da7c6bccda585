"""Mask builders of the model facade (reference: speechbrain/lobes/models/transformer/TransformerASR.py:50-180).

Only the two functions on the SummaryMixing encoder path: the (B,T) padding mask in the SummaryMixing
convention (True = valid, ``masked_false_or_true=False``, :158-162, :348-349) and the (T,T) dynamic-chunk
mask (:85-110).  Both are built on the device by libsmx kernels.
"""
from __future__ import annotations

from typing import Optional

import torch

from .... import _host as H
from .... import _lib as L


def make_transformer_src_mask(src: torch.Tensor, causal: bool = False, masked_false_or_true: bool = True,
                              dynchunktrain_config=None) -> Optional[torch.Tensor]:
    """(T,T) bool mask restricting which frames a frame may summarise over (TransformerASR.py:50-110)."""
    if causal:
        raise NotImplementedError("causal look-ahead masks belong to the self-attention path, not SummaryMixing")
    if dynchunktrain_config is None:
        return None
    H.require_cuda(src, "make_transformer_src_mask")
    T = src.size(1)
    out = torch.empty(T, T, dtype=torch.float32, device=src.device)
    left = dynchunktrain_config.left_context_size
    with torch.cuda.device(src.device):
        L.check(L.lib().smx_chunk_mask(T, int(dynchunktrain_config.chunk_size), -1 if left is None else int(left),
                                       out.data_ptr(), H.stream_ptr(src.device)))
    visible = out != 0
    return ~visible if masked_false_or_true else visible


def make_transformer_src_tgt_masks(src, tgt=None, wav_len=None, pad_idx=0, causal: bool = False,
                                   masked_false_or_true: bool = True, dynchunktrain_config=None):
    """Returns (src_key_padding_mask, None, src_mask, None) for the encoder side (TransformerASR.py:113-180).
    Decoder masks (tgt) are outside the SummaryMixing encoder path."""
    if tgt is not None:
        raise NotImplementedError("decoder masks are outside the SummaryMixing encoder path")
    src_key_padding_mask = None
    if wav_len is not None:
        H.require_cuda(src, "make_transformer_src_tgt_masks")
        B, T = src.shape[0], src.shape[1]
        wl = wav_len.to(device=src.device, dtype=torch.float32).contiguous()
        # the reference's mask width is max(round(wav_len*T)) (length_to_mask); it only works when that is T
        if int(torch.round(wl * T).max().item()) != T:
            raise RuntimeError("padding mask narrower than T: no wav_len entry equals 1.0 (TransformerASR.py:158-162)")
        m = torch.empty(B, T, dtype=torch.uint8, device=src.device)
        with torch.cuda.device(src.device):
            L.check(L.lib().smx_padding_mask_from_wav_len(wl.data_ptr(), B, T, m.data_ptr(), H.stream_ptr(src.device)))
        valid = m != 0
        src_key_padding_mask = ~valid if masked_false_or_true else valid
    src_mask = make_transformer_src_mask(src, causal=causal, masked_false_or_true=masked_false_or_true,
                                         dynchunktrain_config=dynchunktrain_config)
    return src_key_padding_mask, None, src_mask, None
