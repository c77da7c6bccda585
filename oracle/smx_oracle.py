"""CPU oracle for the SummaryMixing hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (plain tensor ops on CPU, fp64 or fp32) of the reference algorithm of
SamsungLabs/SummaryMixing @ d1b1f42 for the path SURVEY.md section 8 scopes.  Each function cites
the reference file:line it follows (paths relative to /root/reference).  Weights are addressed
by the reference's own ``state_dict`` keys, so a reference checkpoint is the oracle's input.

Pinning: the reference's tests hold no numeric assertion (one shape test,
tests/unittests/test_summary_mixing.py:5-57), so this oracle is pinned against outputs of the
reference itself, run in the build container through ``oracle/sbshim`` by ``oracle/gen_golden.py``
and committed under ``tests/golden/`` (``tests/test_oracle_golden.py`` checks every fixture).
Third-party arithmetic that is NOT under /root/reference (SpeechBrain v1.0, unpinned): the
position-wise FFN, LayerNorm wrapper, Swish and the ConvolutionalSpatialGatingUnit are restated
from SpeechBrain's published behaviour — PARITY UNPINNED for those beyond the shim.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package (``summarymixing_b200``) never does.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# activations.  The reference takes an nn.Module class (summary_mixing.py:86,109); the oracle
# takes its name.  Swish: speechbrain.nnet.activations.Swish (x*sigmoid(x)), GELU: exact erf.
# --------------------------------------------------------------------------------------
def activation(name: str, x: Tensor) -> Tensor:
    if name == "swish":
        return x * torch.sigmoid(x)
    if name == "gelu":
        return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    if name == "relu":
        return torch.clamp_min(x, 0.0)
    if name == "leaky_relu":
        return torch.where(x >= 0, x, 0.01 * x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "identity":
        return x
    raise ValueError(f"unknown activation {name}")


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim (biased variance), used at summary_mixing.py:163-165,218,249;
    Conformer.py:135,159,471,479,486-487,759."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _p(sd: SD, key: str, like: Tensor) -> Tensor:
    return sd[key].to(dtype=like.dtype)


# --------------------------------------------------------------------------------------
# VanillaNN / ParallelLinear                                   (VanillaNN.py:26-117, 120-196)
# --------------------------------------------------------------------------------------
def parallel_linear(x: Tensor, weights: Tensor, biases: Tensor, combine_out_dims: bool = True) -> Tensor:
    """Block-diagonal per-head linear, VanillaNN.py:99-117: x viewed (B,T,h,F/h); head m maps its
    F/h slice with weights[m] (F/h, H/h) and adds biases[m]."""
    h, fin, fout = weights.shape
    if x.ndim == 3:
        B, T, Fdim = x.shape
        x = x.reshape(B, T, h, fin)
    out = torch.einsum("btmf,mfh->btmh", x, weights) + biases
    if combine_out_dims:
        out = out.reshape(out.shape[0], out.shape[1], -1)
    return out


def vanilla_nn(x: Tensor, sd: SD, prefix: str, act: str) -> Tensor:
    """VanillaNN.forward (VanillaNN.py:168-196): blocks named linear, linear_0, linear_1, ... each
    followed by the activation — INCLUDING the last block (:196).  n_split>1 blocks are
    ParallelLinear (keys .weights/.biases), else SpeechBrain Linear (keys .w.weight/.w.bias)."""
    names = ["linear"]
    i = 0
    while f"{prefix}linear_{i}.weights" in sd or f"{prefix}linear_{i}.w.weight" in sd:
        names.append(f"linear_{i}")
        i += 1
    for bi, name in enumerate(names):
        last = bi == len(names) - 1
        if f"{prefix}{name}.weights" in sd:
            x = parallel_linear(
                x, _p(sd, f"{prefix}{name}.weights", x), _p(sd, f"{prefix}{name}.biases", x),
                combine_out_dims=last,
            )
        else:
            x = x @ _p(sd, f"{prefix}{name}.w.weight", x).T + _p(sd, f"{prefix}{name}.w.bias", x)
        x = activation(act, x)
    return x


# --------------------------------------------------------------------------------------
# SummaryMixing cell                                                (summary_mixing.py:169-379)
# --------------------------------------------------------------------------------------
def laplace_weights(T: int, decay: Tensor, binary_mask: Optional[Tensor], dtype) -> Tensor:
    """summary_mixing.py:330-379 with normalise=False: decay**|i-j| (* binary mask)."""
    idx = torch.arange(T)
    dist = (idx[None, :] - idx[:, None]).abs().to(dtype)
    w = torch.exp(dist * torch.log(decay.to(dtype)))
    if binary_mask is not None:
        w = w * binary_mask.to(dtype)
    return w


def summary_mixing(
    x: Tensor,
    sd: SD,
    prefix: str = "",
    mode: str = "SummaryMixing",
    act: str = "gelu",
    use_layernorm: bool = True,
    src_padding_mask: Optional[Tensor] = None,
    sum_mask: Optional[Tensor] = None,
    drop=None,
) -> Tensor:
    """SummaryMixing.forward; eval mode (dropout = identity) unless ``drop`` is given: drop(key, tensor) stands in for
    self.dropout on the concatenation (:252, :297), key = prefix + "cat" (oracle/dropout.py).

    mask prep summary_mixing.py:183-189; full/expdecay :198-253; fast :255-298; lite :300-324.
    ``src_padding_mask`` (B,T): 1/True = valid frame (TransformerASR.py:159-162,348-349)."""
    B, T, _ = x.shape
    if src_padding_mask is None:
        m = torch.ones(B, T, 1, dtype=x.dtype)
    else:
        m = src_padding_mask.to(x.dtype).unsqueeze(-1)
    if sum_mask is not None:
        sum_mask = sum_mask.to(x.dtype)  # .float() at :189

    if mode in ("SummaryMixing", "SummaryMixing-expdecay"):
        local = vanilla_nn(x, sd, prefix + "local_proj.", act) * m  # :215
        if use_layernorm:
            local = layer_norm(local, _p(sd, prefix + "local_norm.weight", x), _p(sd, prefix + "local_norm.bias", x))
        s = vanilla_nn(x, sd, prefix + "summary_proj.", act) * m  # :221
        if mode == "SummaryMixing-expdecay":
            sum_mask = laplace_weights(T, sd[prefix + "decay_constant"], sum_mask, x.dtype)  # :223-224
        if sum_mask is None:
            summ = s.sum(dim=1) / m.sum(dim=1)  # :229-231
            summ = summ.unsqueeze(1).repeat(1, T, 1)  # :233
        else:
            summ = torch.matmul(sum_mask, s) / sum_mask.sum(dim=1).unsqueeze(-1)  # :244-246
        if use_layernorm:
            summ = layer_norm(summ, _p(sd, prefix + "summary_norm.weight", x), _p(sd, prefix + "summary_norm.bias", x))
        cat = torch.cat([local, summ], dim=-1)  # :251-253
        if drop is not None:
            cat = drop(prefix + "cat", cat)  # :252
        return vanilla_nn(cat, sd, prefix + "summary_local_merging.", act)

    if mode == "SummaryMixing-fast":
        g = vanilla_nn(x, sd, prefix + "global_proj.", act) * m  # :271
        d_l = g.shape[-1] // 2
        local, s = g[..., :d_l], g[..., d_l:]  # :272-275 (split in chunks of local_proj_out_dim)
        if sum_mask is None:
            summ = s.sum(dim=1) / m.sum(dim=1)  # :278-280
            summ = summ.unsqueeze(1).repeat(1, T, 1)
        else:
            summ = torch.matmul(sum_mask, s) / sum_mask.sum(dim=1).unsqueeze(-1)  # :292-294
        cat = torch.cat([local, summ], dim=-1)
        if drop is not None:
            cat = drop(prefix + "cat", cat)  # :297
        return vanilla_nn(cat, sd, prefix + "summary_local_merging.", act)  # :296-298

    if mode == "SummaryMixing-lite":
        s = vanilla_nn(x, sd, prefix + "summary_proj.", act) * m  # :318
        summ = s.sum(dim=1) / m.sum(dim=1)  # :319-321
        return summ.unsqueeze(1).expand(-1, T, -1)  # :322 (no LN, no combiner)

    raise ValueError(
        "The SummaryMixing mode should either be 'SummaryMixing', 'SummaryMixing-lite', "
        "'SummaryMixing-fast' or 'SummaryMixing-expdecay'"
    )


# --------------------------------------------------------------------------------------
# Conformer                                                                 (Conformer.py)
# --------------------------------------------------------------------------------------
def convolution_module(
    x: Tensor,
    sd: SD,
    prefix: str,
    act: str = "swish",
    mask: Optional[Tensor] = None,
    causal: bool = False,
    masked_false_or_true: bool = False,
    chunk_size: Optional[int] = None,
    drop=None,
) -> Tensor:
    """ConvolutionModule.forward, Conformer.py:166-340 (dilation 1); ``drop`` (training mode): the nn.Dropout ending after_conv (:163), key prefix + "out".  Non-chunked :322-332, chunked
    (Dynamic Chunk Convolution) :197-320, output masking :334-338.  ``mask`` is (B,T,1)."""
    cw = _p(sd, prefix + "conv.weight", x)  # (D,1,k)
    D, _, k = cw.shape
    pad = (k - 1) if causal else (k - 1) // 2  # :130-133
    out = layer_norm(x, _p(sd, prefix + "layer_norm.weight", x), _p(sd, prefix + "layer_norm.bias", x))
    out = out.transpose(1, 2)
    out = F.conv1d(out, _p(sd, prefix + "bottleneck.0.weight", x), _p(sd, prefix + "bottleneck.0.bias", x))
    a, g = out.chunk(2, dim=1)
    out = a * torch.sigmoid(g)  # nn.GLU(dim=1) :139
    cb = _p(sd, prefix + "conv.bias", x)
    if chunk_size is not None:
        assert not causal
        B, _, T = out.shape
        frp = (chunk_size - T % chunk_size) % chunk_size  # :222-225
        out = F.pad(out, (pad, frp))  # :237
        out = out.unfold(2, size=chunk_size + pad, step=chunk_size)  # :256
        out = F.pad(out, (0, pad))  # :266
        out = out.transpose(1, 2).flatten(0, 1)  # :273-276
        out = F.conv1d(out, cw, cb, padding=0, groups=D)  # :297-305
        out = out.transpose(1, 2)
    else:
        out = F.conv1d(out, cw, cb, padding=pad, groups=D)  # :325
        if causal:
            out = out[..., :-pad]  # :327-329
        out = out.transpose(1, 2)
    out = layer_norm(out, _p(sd, prefix + "after_conv.0.weight", x), _p(sd, prefix + "after_conv.0.bias", x))
    out = activation(act, out)
    out = out @ _p(sd, prefix + "after_conv.2.weight", x).T + _p(sd, prefix + "after_conv.2.bias", x)
    if drop is not None:
        out = drop(prefix + "out", out)  # after_conv[3]
    if chunk_size is not None:
        out = out.reshape(B, -1, D)  # :313-316
        if frp > 0:
            out = out[:, :-frp, :]
    if mask is not None:
        if masked_false_or_true:
            out = out.masked_fill(mask.bool(), 0.0)  # :336
        else:
            out = out * mask.to(out.dtype)  # :338
    return out


def ffn_module(x: Tensor, sd: SD, prefix: str, act: str, drop=None) -> Tensor:
    """ffn_module{1,2} = Sequential(LayerNorm, PositionalwiseFeedForward, Dropout), Conformer.py:470-484;
    PWFF (SpeechBrain, unpinned) = Linear(D,d_ffn) -> act -> Dropout -> Linear(d_ffn,D)."""
    h = layer_norm(x, _p(sd, prefix + "0.weight", x), _p(sd, prefix + "0.bias", x))
    h = h @ _p(sd, prefix + "1.ffn.0.weight", x).T + _p(sd, prefix + "1.ffn.0.bias", x)
    h = activation(act, h)
    if drop is not None:
        h = drop(prefix + "inner", h)  # PositionalwiseFeedForward's dropout
    out = h @ _p(sd, prefix + "1.ffn.3.weight", x).T + _p(sd, prefix + "1.ffn.3.bias", x)
    return drop(prefix + "outer", out) if drop is not None else out  # ffn_module[2]


def conformer_layer(
    x: Tensor,
    sd: SD,
    prefix: str,
    act: str = "swish",
    mode: str = "SummaryMixing",
    use_layernorm: bool = True,
    src_mask: Optional[Tensor] = None,
    src_key_padding_mask: Optional[Tensor] = None,
    causal: bool = False,
    chunk_size: Optional[int] = None,
    drop=None,
) -> Tensor:
    """ConformerEncoderLayer.forward with attention_type == 'SummaryMixing', Conformer.py:490-548 (``drop``: training mode)."""
    conv_mask = None if src_key_padding_mask is None else src_key_padding_mask.unsqueeze(-1)  # :514-516
    x = x + 0.5 * ffn_module(x, sd, prefix + "ffn_module1.", act, drop)  # :518
    skip = x
    x = layer_norm(x, _p(sd, prefix + "norm1.norm.weight", x), _p(sd, prefix + "norm1.norm.bias", x))  # :521
    x = summary_mixing(
        x, sd, prefix + "mha_layer.", mode=mode, act=act, use_layernorm=use_layernorm,
        src_padding_mask=src_key_padding_mask, sum_mask=src_mask, drop=drop,
    )  # :524-526
    x = x + skip  # :541
    x = x + convolution_module(
        x, sd, prefix + "convolution_module.", act=act, mask=conv_mask, causal=causal,
        masked_false_or_true=False, chunk_size=chunk_size, drop=drop,
    )  # :543-545
    y = x + 0.5 * ffn_module(x, sd, prefix + "ffn_module2.", act, drop)
    return layer_norm(y, _p(sd, prefix + "norm2.norm.weight", x), _p(sd, prefix + "norm2.norm.bias", x))  # :547


def conformer_encoder(
    x: Tensor,
    sd: SD,
    num_layers: int,
    prefix: str = "",
    act: str = "swish",
    mode: str = "SummaryMixing",
    use_layernorm: bool = True,
    src_mask: Optional[Tensor] = None,
    src_key_padding_mask: Optional[Tensor] = None,
    causal: bool = False,
    chunk_size: Optional[int] = None,
    drop=None,
) -> Tensor:
    """ConformerEncoder.forward (no layerdrop; ``drop``: training-mode dropout hook), Conformer.py:797-827; final LN eps=1e-6 (:759)."""
    for i in range(num_layers):
        x = conformer_layer(
            x, sd, f"{prefix}layers.{i}.", act=act, mode=mode, use_layernorm=use_layernorm,
            src_mask=src_mask, src_key_padding_mask=src_key_padding_mask, causal=causal, chunk_size=chunk_size, drop=drop,
        )
    return layer_norm(x, _p(sd, prefix + "norm.norm.weight", x), _p(sd, prefix + "norm.norm.bias", x), eps=1e-6)


# --------------------------------------------------------------------------------------
# Self-attention Conformer (the comparison arm of BASELINE.json configs[4])   Conformer.py:425-429, 528-541
# --------------------------------------------------------------------------------------
def conformer_layer_mhsa(x: Tensor, sd: SD, prefix: str, nhead: int, act: str = "swish",
                         key_padding_mask: Optional[Tensor] = None) -> Tensor:
    """ConformerEncoderLayer.forward with attention_type == 'regularMHA' (Conformer.py:425-429, 518-548): the same layer with
    SpeechBrain's MultiheadAttention (a wrapper of nn.MultiheadAttention: packed in_proj, out_proj; un-vendored) in place of
    the SummaryMixing cell.  key_padding_mask (B,T): True = PADDED frame (the MHSA convention, TransformerASR.py:159-162); the
    convolution module then zeroes padded frames with masked_fill (masked_false_or_true=True, Conformer.py:334-338)."""
    B, T, D = x.shape
    x = x + 0.5 * ffn_module(x, sd, prefix + "ffn_module1.", act)
    skip = x
    xn = layer_norm(x, _p(sd, prefix + "norm1.norm.weight", x), _p(sd, prefix + "norm1.norm.bias", x))
    qkv = xn @ _p(sd, prefix + "mha_layer.att.in_proj_weight", x).T + _p(sd, prefix + "mha_layer.att.in_proj_bias", x)
    q, k, v = (t.reshape(B, T, nhead, D // nhead).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    am = None if key_padding_mask is None else (~key_padding_mask)[:, None, None, :]  # True = may attend
    a = F.scaled_dot_product_attention(q, k, v, attn_mask=am)
    a = a.transpose(1, 2).reshape(B, T, D)
    x = a @ _p(sd, prefix + "mha_layer.att.out_proj.weight", x).T + _p(sd, prefix + "mha_layer.att.out_proj.bias", x) + skip
    conv_mask = None if key_padding_mask is None else key_padding_mask.unsqueeze(-1)
    x = x + convolution_module(x, sd, prefix + "convolution_module.", act=act, mask=conv_mask, masked_false_or_true=True)
    y = x + 0.5 * ffn_module(x, sd, prefix + "ffn_module2.", act)
    return layer_norm(y, _p(sd, prefix + "norm2.norm.weight", x), _p(sd, prefix + "norm2.norm.bias", x))


def conformer_encoder_mhsa(x: Tensor, sd: SD, num_layers: int, nhead: int, prefix: str = "", act: str = "swish",
                           key_padding_mask: Optional[Tensor] = None) -> Tensor:
    """ConformerEncoder.forward (Conformer.py:797-827) over conformer_layer_mhsa layers; final LN eps=1e-6."""
    for i in range(num_layers):
        x = conformer_layer_mhsa(x, sd, f"{prefix}layers.{i}.", nhead, act=act, key_padding_mask=key_padding_mask)
    return layer_norm(x, _p(sd, prefix + "norm.norm.weight", x), _p(sd, prefix + "norm.norm.bias", x), eps=1e-6)


# --------------------------------------------------------------------------------------
# Branchformer                                                            (Branchformer.py)
# --------------------------------------------------------------------------------------
def csgu(x: Tensor, sd: SD, prefix: str, gate_act: str = "identity", drop=None) -> Tensor:
    """SpeechBrain ConvolutionalSpatialGatingUnit (UNPINNED, see module docstring): split halves,
    LN the gate half, depthwise conv 'same' with reflect padding, optional linear, gate act, multiply, dropout
    (``drop``: training mode, key prefix + "out")."""
    x1, x2 = x.chunk(2, dim=-1)
    x2 = layer_norm(x2, _p(sd, prefix + "norm.norm.weight", x), _p(sd, prefix + "norm.norm.bias", x))
    cw = _p(sd, prefix + "conv.conv.weight", x)
    C, _, k = cw.shape
    pad = (k - 1) // 2
    x2 = F.pad(x2.transpose(1, 2), (pad, pad), mode="reflect")
    x2 = F.conv1d(x2, cw, _p(sd, prefix + "conv.conv.bias", x), groups=C).transpose(1, 2)
    if prefix + "linear.weight" in sd:
        x2 = x2 @ _p(sd, prefix + "linear.weight", x).T + _p(sd, prefix + "linear.bias", x)
    x2 = activation(gate_act, x2)
    out = x2 * x1
    return drop(prefix + "out", out) if drop is not None else out


def convolution_branch(x: Tensor, sd: SD, prefix: str, act: str = "gelu", gate_act: str = "identity", drop=None) -> Tensor:
    """ConvolutionBranch.forward, Branchformer.py:86-97 (``drop``: the CSGU's dropout in training mode)."""
    x = activation(act, x @ _p(sd, prefix + "pre_channel_proj.weight", x).T + _p(sd, prefix + "pre_channel_proj.bias", x))
    x = csgu(x, sd, prefix + "csgu.", gate_act, drop)
    return x @ _p(sd, prefix + "post_channel_proj.weight", x).T + _p(sd, prefix + "post_channel_proj.bias", x)


def branchformer_layer(
    x: Tensor,
    sd: SD,
    prefix: str,
    act: str = "gelu",
    gate_act: str = "identity",
    mode: str = "SummaryMixing",
    src_mask: Optional[Tensor] = None,
    src_key_padding_mask: Optional[Tensor] = None,
    drop=None,
) -> Tensor:
    """BranchformerEncoderLayer.forward with attention_type == 'SummaryMixing', Branchformer.py:243-334.  ``drop`` (training mode):
    drop(key, tensor) stands in for self.dropout after each branch (keys prefix + "x1", prefix + "x2": :334, :294), after merge_proj
    (prefix + "merge": :279) and for the CSGU's own dropout (prefix + "convolution_branch.csgu.out"); the cell's dropout
    (prefix + "mha_layer.cat") as in summary_mixing()."""
    x1 = layer_norm(x, _p(sd, prefix + "norm_mhsa.norm.weight", x), _p(sd, prefix + "norm_mhsa.norm.bias", x))  # :317
    x1 = summary_mixing(
        x1, sd, prefix + "mha_layer.", mode=mode, act=act, use_layernorm=True,
        src_padding_mask=src_key_padding_mask, sum_mask=src_mask, drop=drop if mode != "SummaryMixing-lite" else None,
    )  # :320-322
    x2 = layer_norm(x, _p(sd, prefix + "norm_conv.norm.weight", x), _p(sd, prefix + "norm_conv.norm.bias", x))  # :292
    x2 = convolution_branch(x2, sd, prefix + "convolution_branch.", act=act, gate_act=gate_act, drop=drop)  # :293 (no mask, :276)
    if drop is not None:
        x1 = drop(prefix + "x1", x1.contiguous())  # :334
        x2 = drop(prefix + "x2", x2)  # :294
    merged = vanilla_nn(torch.cat([x1, x2], dim=-1), sd, prefix + "merge_proj.", act)  # :279, :220-226
    return x + (drop(prefix + "merge", merged) if drop is not None else merged)


def branchformer_encoder(
    x: Tensor,
    sd: SD,
    num_layers: int,
    prefix: str = "",
    act: str = "gelu",
    gate_act: str = "identity",
    mode: str = "SummaryMixing",
    src_mask: Optional[Tensor] = None,
    src_key_padding_mask: Optional[Tensor] = None,
    drop=None,
) -> Tensor:
    """BranchformerEncoder.forward, Branchformer.py:479-491; final LN eps=1e-6 (:444); ``drop``: training-mode dropout hook."""
    for i in range(num_layers):
        x = branchformer_layer(
            x, sd, f"{prefix}layers.{i}.", act=act, gate_act=gate_act, mode=mode,
            src_mask=src_mask, src_key_padding_mask=src_key_padding_mask, drop=drop,
        )
    return layer_norm(x, _p(sd, prefix + "norm.norm.weight", x), _p(sd, prefix + "norm.norm.bias", x), eps=1e-6)


# --------------------------------------------------------------------------------------
# mask builders                                                      (TransformerASR.py:50-180)
# --------------------------------------------------------------------------------------
def padding_mask_from_wav_len(wav_len: Tensor, T: int) -> Tensor:
    """make_transformer_src_tgt_masks with masked_false_or_true=False, TransformerASR.py:157-162:
    abs_len = round(wav_len * T); mask = arange(max(abs_len)) < abs_len (True = valid).  The
    reference's mask width is max(abs_len), so T must equal it."""
    abs_len = torch.round(wav_len * T)
    width = int(abs_len.max().item())
    if width != T:
        raise RuntimeError(f"padding mask width {width} != T {T} (no wav_len entry equals 1.0)")
    return torch.arange(T)[None, :] < abs_len[:, None]


def chunk_mask(T: int, chunk_size: int, left_context_size: Optional[int] = None) -> Tensor:
    """make_transformer_src_mask with masked_false_or_true=False, TransformerASR.py:85-110:
    frame t sees frames [chunk_start - left*chunk, chunk_end) (True = visible)."""
    t = torch.arange(T)
    end = (t // chunk_size + 1) * chunk_size
    m = t[None, :] < end[:, None]
    if left_context_size is not None:
        start = end - chunk_size * (left_context_size + 1)
        m = m & (t[None, :] >= start[:, None])
    return m


def sinusoidal_positional_encoding(T: int, D: int, dtype=torch.float32) -> Tensor:
    """PositionalEncoding, Transformer.py:288-339 (table built in fp32 there)."""
    pe = torch.zeros(T, D, dtype=torch.float32)
    pos = torch.arange(0, T).unsqueeze(1).float()
    den = torch.exp(torch.arange(0, D, 2).float() * -(math.log(10000.0) / D))
    pe[:, 0::2] = torch.sin(pos * den)
    pe[:, 1::2] = torch.cos(pos * den)
    return pe.unsqueeze(0).to(dtype)
