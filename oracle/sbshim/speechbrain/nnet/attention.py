"""speechbrain.nnet.attention stand-in.

``PositionalwiseFeedForward``: Linear(in,d_ffn) -> activation -> Dropout -> Linear(d_ffn,in) held as
``self.ffn`` (indices 0 and 3 carry weights); upstream permutes (B,T,D)->(T,B,D) and back around it.
``MultiheadAttention`` wraps torch.nn.MultiheadAttention as ``self.att`` (needed for the decoder and
the self-attention comparison).  RelPos* are import-only stubs.
"""
from typing import Optional

import torch
import torch.nn as nn


class PositionalwiseFeedForward(nn.Module):
    def __init__(self, d_ffn, input_shape=None, input_size=None, dropout=0.0, activation=nn.ReLU):
        super().__init__()
        if input_shape is None and input_size is None:
            raise ValueError("Expected one of input_shape or input_size")
        if input_size is None:
            input_size = input_shape[-1]
        self.ffn = nn.Sequential(
            nn.Linear(input_size, d_ffn),
            activation(),
            nn.Dropout(dropout),
            nn.Linear(d_ffn, input_size),
        )

    def forward(self, x):
        x = x.permute(1, 0, 2)
        x = self.ffn(x)
        x = x.permute(1, 0, 2)
        return x


class MultiheadAttention(nn.Module):
    def __init__(self, nhead, d_model, dropout=0.0, bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None, vdim=None):
        super().__init__()
        self.att = nn.MultiheadAttention(
            embed_dim=d_model, num_heads=nhead, dropout=dropout, bias=bias,
            add_bias_kv=add_bias_kv, add_zero_attn=add_zero_attn, kdim=kdim, vdim=vdim,
        )

    def forward(self, query, key, value, attn_mask: Optional[torch.Tensor] = None,
                key_padding_mask: Optional[torch.Tensor] = None, return_attn_weights: bool = True,
                pos_embs: Optional[torch.Tensor] = None):
        query = query.permute(1, 0, 2)
        key = key.permute(1, 0, 2)
        value = value.permute(1, 0, 2)
        if attn_mask is not None and key_padding_mask is not None and attn_mask.dtype != key_padding_mask.dtype:
            key_padding_mask = key_padding_mask.to(attn_mask.dtype)
        output, attention_weights = self.att(
            query, key, value, attn_mask=attn_mask, key_padding_mask=key_padding_mask,
            need_weights=return_attn_weights,
        )
        output = output.permute(1, 0, 2)
        if return_attn_weights:
            return output, attention_weights
        return output


class RelPosMHAXL(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("RelPosMHAXL is outside the SummaryMixing hot path (SURVEY.md section 8c)")


class RelPosEncXL(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("RelPosEncXL is outside the SummaryMixing hot path (SURVEY.md section 8c)")
