"""Golden GRADIENTS of the SummaryMixing cell, the convolution module, the Conformer layer / encoder and the Branchformer encoder from the UNMODIFIED reference (run in the build container).

    python oracle/gen_golden_bwd.py        # writes tests/golden/bwd/*.npz

For each listed forward fixture (tests/golden/<name>.npz, written by gen_golden.py) the reference module
(speechbrain/nnet/summary_mixing.py, imported from /root/reference through oracle/sbshim) is rebuilt from the fixture's
config and state_dict, run in eval mode (dropout off) with autograd on the fixture's input, and back-propagated from a
seeded dy.  Stored: dy, dx and one gradient per state_dict key.  tests/ compare the oracle's autograd (CPU) and
smx_summary_mixing_bwd (GPU) against these.  No reference source is copied; test infrastructure only.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "sbshim"))
sys.path.insert(0, ROOT)

from speechbrain.nnet.activations import Swish  # noqa: E402
from speechbrain.nnet.summary_mixing import SummaryMixing  # noqa: E402
from speechbrain.lobes.models.VanillaNN import VanillaNN  # noqa: E402
from speechbrain.lobes.models.transformer.Branchformer import BranchformerEncoder  # noqa: E402
from speechbrain.lobes.models.transformer.Conformer import (  # noqa: E402
    ConformerEncoder,
    ConformerEncoderLayer,
    ConvolutionModule,
)
from speechbrain.utils.dynamic_chunk_training import DynChunkTrainConfig  # noqa: E402
from tests import _golden as G  # noqa: E402

ACTS = {"swish": Swish, "gelu": nn.GELU, "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU}
CASES = ["cell_sm_h4_swish", "cell_sm_h1_gelu", "cell_sm_h4_relu_noln_deep_nomask", "cell_reftest_sm_h4", "cell_sm_h4_gelu", "cell_sm_lite_h4_gelu", "cell_sm_lite_h1_gelu",
         "convmod_plain", "convmod_causal", "conformer_layer", "conformer_enc_sm_h4", "conformer_enc_sm_h1_gelu",
         "cell_sm_fast_h4_swish", "cell_sm_fast_h1_gelu", "conformer_enc_lite_h4", "vanilla_split1", "vanilla_split3",
         "branchformer_enc_lite", "branchformer_enc_full",
         "cell_sm_h4_swish_summask", "cell_sm_fast_h4_swish_summask", "convmod_dcconv", "conformer_enc_fast_noln_dynchunk",
         "cell_sm_expdecay_h4_gelu", "cell_sm_expdecay_h1_gelu"]
OUT = os.path.join(ROOT, "tests", "golden", "bwd")


def main():
    os.makedirs(OUT, exist_ok=True)
    for i, name in enumerate(CASES):
        if "--all" not in sys.argv and os.path.exists(os.path.join(OUT, name + ".npz")):
            continue
        fx = G.Fixture(name)
        c = fx.cfg
        k = c["kind"]
        if k == "cell":
            sm = SummaryMixing(c["enc_dim"], c["nhead"], c["local_proj_hid_dim"], c["local_proj_out_dim"], c["summary_hid_dim"],
                               c["summary_out_dim"], activation=ACTS[c["act"]], mode=c["mode"], use_layernorm=c["use_layernorm"])
            run = (lambda m, x: m(x, sum_mask=fx.sum_mask, src_padding_mask=fx.mask)) if fx.mask is not None else (lambda m, x: m(x))
        elif k == "vanilla":
            sm = VanillaNN(input_shape=[None, None, c["input_size"]], activation=ACTS[c["act"]], dnn_blocks=len(c["dnn_neurons"]),
                           dnn_neurons=c["dnn_neurons"], n_split=c["n_split"])
            run = lambda m, x: m(x)  # noqa: E731
        elif k == "conv_module":
            sm = ConvolutionModule(c["input_size"], c["kernel_size"], True, ACTS[c["act"]], 0.0, causal=c["causal"],
                                   masked_false_or_true=False)
            dcc = DynChunkTrainConfig(c["chunk_size"]) if c.get("chunk_size") else None
            run = lambda m, x: m(x, fx.mask.unsqueeze(-1), dynchunktrain_config=dcc)  # noqa: E731
        elif k == "conformer_layer":
            sm = ConformerEncoderLayer(c["d_model"], c["d_ffn"], c["nhead"], c["kernel_size"], activation=ACTS[c["act"]],
                                       attention_type="SummaryMixing", local_proj_hid_dim=c["local_proj_hid_dim"],
                                       local_proj_out_dim=c["local_proj_out_dim"], summary_hid_dim=c["summary_hid_dim"],
                                       mode=c["mode"], use_layernorm=c["use_layernorm"])
            run = lambda m, x: m(x, src_key_padding_mask=fx.mask)[0]  # noqa: E731
        elif k == "branchformer_encoder":
            sm = BranchformerEncoder(c["num_layers"], c["d_model"], c["nhead"], c["kernel_size"], csgu_linear_units=c["csgu_linear_units"],
                                     local_proj_hid_dim=c["local_proj_hid_dim"], local_proj_out_dim=c["local_proj_out_dim"],
                                     summary_hid_dim=c["summary_hid_dim"], summary_out_dim=c["summary_out_dim"], mode=c["mode"])
            run = lambda m, x: m(x, src_key_padding_mask=fx.mask)[0]  # noqa: E731
        else:
            sm = ConformerEncoder(c["num_layers"], c["d_model"], c["d_ffn"], c["nhead"], c["kernel_size"],
                                  activation=ACTS[c["act"]], attention_type="SummaryMixing",
                                  local_proj_hid_dim=c["local_proj_hid_dim"], local_proj_out_dim=c["local_proj_out_dim"],
                                  summary_hid_dim=c["summary_hid_dim"], mode=c["mode"], use_layernorm=c["use_layernorm"])
            dce = DynChunkTrainConfig(c["chunk_size"]) if c.get("chunk_size") else None
            run = lambda m, x: m(x, src_mask=fx.sum_mask, src_key_padding_mask=fx.mask, dynchunktrain_config=dce)[0]  # noqa: E731
        sm.load_state_dict(fx.sd)
        sm.eval()
        x = fx.x.clone().requires_grad_(True)
        y = run(sm, x)
        assert float((y.detach() - fx.y).abs().max()) < 1e-6, name
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(7000 + i))
        y.backward(dy)
        arrays = {"dy": dy.numpy(), "dx": x.grad.numpy()}
        for k, p in sm.named_parameters():
            if p.grad is not None:
                arrays["grad." + k] = p.grad.numpy()
        arrays["cfg"] = np.frombuffer(json.dumps(dict(forward_fixture=name, torch=torch.__version__)).encode(), dtype=np.uint8)
        np.savez(os.path.join(OUT, name + ".npz"), **arrays)
        print(name, "dx max", float(x.grad.abs().max()), "grads", len(arrays) - 3)


if __name__ == "__main__":
    main()
