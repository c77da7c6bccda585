"""Build product modules (summarymixing_b200) from a golden fixture's config and load its state_dict."""
import torch.nn as nn

import summarymixing_b200 as S

ACTS = {"swish": S.Swish, "gelu": nn.GELU, "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "identity": nn.Identity}


def module_from_fixture(fx):
    c = fx.cfg
    k = c["kind"]
    if k == "cell":
        m = S.SummaryMixing(c["enc_dim"], c["nhead"], c["local_proj_hid_dim"], c["local_proj_out_dim"],
                            c["summary_hid_dim"], c["summary_out_dim"], activation=ACTS[c["act"]], mode=c["mode"],
                            use_layernorm=c["use_layernorm"])
    elif k == "vanilla":
        m = S.VanillaNN(input_shape=[None, None, c["input_size"]], activation=ACTS[c["act"]],
                        dnn_blocks=len(c["dnn_neurons"]), dnn_neurons=c["dnn_neurons"], n_split=c["n_split"])
    elif k == "conv_module":
        m = S.ConvolutionModule(c["input_size"], c["kernel_size"], True, ACTS[c["act"]], 0.0, causal=c["causal"],
                                masked_false_or_true=False)
    elif k == "conformer_layer":
        m = S.ConformerEncoderLayer(c["d_model"], c["d_ffn"], c["nhead"], c["kernel_size"], activation=ACTS[c["act"]],
                                    attention_type="SummaryMixing", local_proj_hid_dim=c["local_proj_hid_dim"],
                                    local_proj_out_dim=c["local_proj_out_dim"], summary_hid_dim=c["summary_hid_dim"],
                                    mode=c["mode"], use_layernorm=c["use_layernorm"])
    elif k == "conformer_encoder":
        m = S.ConformerEncoder(c["num_layers"], c["d_model"], c["d_ffn"], c["nhead"], c["kernel_size"],
                               activation=ACTS[c["act"]], attention_type="SummaryMixing",
                               local_proj_hid_dim=c["local_proj_hid_dim"], local_proj_out_dim=c["local_proj_out_dim"],
                               summary_hid_dim=c["summary_hid_dim"], mode=c["mode"], use_layernorm=c["use_layernorm"])
    elif k == "branchformer_encoder":
        m = S.BranchformerEncoder(c["num_layers"], c["d_model"], c["nhead"], c["kernel_size"],
                                  activation=ACTS[c["act"]], csgu_linear_units=c["csgu_linear_units"],
                                  gate_activation=ACTS[c["gate_act"]], local_proj_hid_dim=c["local_proj_hid_dim"],
                                  local_proj_out_dim=c["local_proj_out_dim"], summary_hid_dim=c["summary_hid_dim"],
                                  summary_out_dim=c["summary_out_dim"], mode=c["mode"])
    else:
        raise AssertionError(k)
    m.load_state_dict(fx.sd, strict=True)
    return m.eval()


class _DC:
    def __init__(self, chunk_size, left_context_size=None):
        self.chunk_size = chunk_size
        self.left_context_size = left_context_size


def run_module(m, fx, x, device):
    """Call the product module the way the reference's caller does; returns the output tensor."""
    c = fx.cfg
    k = c["kind"]
    mask = None if fx.mask is None else fx.mask.to(device)
    smask = None if fx.sum_mask is None else fx.sum_mask.to(device)
    if k == "cell":
        return m(x, sum_mask=smask, src_padding_mask=mask)
    if k == "vanilla":
        return m(x)
    if k == "conv_module":
        dc = None if c["chunk_size"] is None else _DC(c["chunk_size"])
        return m(x, mask.unsqueeze(-1), dynchunktrain_config=dc)
    if k == "conformer_layer":
        return m(x, src_key_padding_mask=mask)[0]
    if k == "conformer_encoder":
        dc = None if c["chunk_size"] is None else _DC(c["chunk_size"])
        return m(x, src_mask=smask, src_key_padding_mask=mask, dynchunktrain_config=dc)[0]
    if k == "branchformer_encoder":
        return m(x, src_key_padding_mask=mask)[0]
    raise AssertionError(k)
