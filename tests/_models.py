"""Builders shared by the tile-aligned / full-size parity tests: this repo's modules and the oracle call for a fixture cfg."""
import torch.nn as nn

import summarymixing_b200 as S
from oracle import smx_oracle as O
from oracle.seeded import fill_module

ACTS = {"swish": S.Swish, "gelu": nn.GELU, "relu": nn.ReLU}


def build(cfg):
    """This repo's module for a fixture config, with the fixture's seed-defined weights."""
    k = cfg["kind"]
    if k == "cell":
        m = S.SummaryMixing(cfg["enc_dim"], cfg["nhead"], cfg["local_proj_hid_dim"], cfg["local_proj_out_dim"], cfg["summary_hid_dim"],
                            cfg["summary_out_dim"], activation=ACTS[cfg["act"]], mode=cfg["mode"], use_layernorm=cfg["use_layernorm"])
    elif k == "conv_module":
        m = S.ConvolutionModule(cfg["input_size"], cfg["kernel_size"], True, ACTS[cfg["act"]], 0.0, masked_false_or_true=False)
    elif k == "conformer_layer":
        m = S.ConformerEncoderLayer(cfg["d_model"], cfg["d_ffn"], cfg["nhead"], cfg["kernel_size"], activation=ACTS[cfg["act"]],
                                    attention_type="SummaryMixing", local_proj_hid_dim=cfg["local_proj_hid_dim"],
                                    local_proj_out_dim=cfg["local_proj_out_dim"], summary_hid_dim=cfg["summary_hid_dim"],
                                    mode=cfg["mode"], use_layernorm=cfg["use_layernorm"])
    elif k == "conformer_encoder":
        m = S.ConformerEncoder(cfg["num_layers"], cfg["d_model"], cfg["d_ffn"], cfg["nhead"], cfg["kernel_size"],
                               activation=ACTS[cfg["act"]], attention_type="SummaryMixing", local_proj_hid_dim=cfg["local_proj_hid_dim"],
                               local_proj_out_dim=cfg["local_proj_out_dim"], summary_hid_dim=cfg["summary_hid_dim"], mode=cfg["mode"],
                               use_layernorm=cfg["use_layernorm"])
    elif k == "branchformer_encoder":
        m = S.BranchformerEncoder(cfg["num_layers"], cfg["d_model"], cfg["nhead"], cfg["kernel_size"],
                                  csgu_linear_units=cfg["csgu_linear_units"], local_proj_hid_dim=cfg["local_proj_hid_dim"],
                                  local_proj_out_dim=cfg["local_proj_out_dim"], summary_hid_dim=cfg["summary_hid_dim"],
                                  summary_out_dim=cfg["summary_out_dim"], mode=cfg["mode"])
    else:
        raise ValueError(k)
    fill_module(m, cfg["seed_w"])
    return m.eval()


def run_module(m, cfg, x, mask):
    """Forward of this repo's module in the fixture's calling convention; returns a (B,T,D) tensor."""
    k = cfg["kind"]
    if k == "cell":
        return m(x, src_padding_mask=mask).contiguous()
    if k == "conv_module":
        return m(x, mask.unsqueeze(-1))
    return m(x, src_key_padding_mask=mask)[0]


def run_oracle(cfg, sd, x, mask):
    """The CPU oracle on the same state_dict (any dtype: fp32, or bf16 for the reference algorithm's own bf16 error)."""
    k = cfg["kind"]
    if k == "cell":
        y = O.summary_mixing(x, sd, mode=cfg["mode"], act=cfg["act"], src_padding_mask=mask, use_layernorm=cfg["use_layernorm"])
        return y.contiguous()
    if k == "conv_module":
        return O.convolution_module(x, sd, "", act=cfg["act"], mask=mask.unsqueeze(-1))
    if k == "conformer_layer":
        return O.conformer_layer(x, sd, "", act=cfg["act"], src_key_padding_mask=mask, mode=cfg["mode"], use_layernorm=cfg["use_layernorm"])
    if k == "conformer_encoder":
        return O.conformer_encoder(x, sd, cfg["num_layers"], act=cfg["act"], src_key_padding_mask=mask, mode=cfg["mode"],
                                   use_layernorm=cfg["use_layernorm"])
    if k == "branchformer_encoder":
        return O.branchformer_encoder(x, sd, cfg["num_layers"], act=cfg["act"], gate_act=cfg["gate_act"], src_key_padding_mask=mask,
                                      mode=cfg["mode"])
    raise ValueError(k)
