"""Pin the CPU oracle (oracle/smx_oracle.py) against outputs of the unmodified reference.

The reference's own tests hold no numeric assertion, so the golden vectors under
tests/golden/ — produced by running the reference itself (oracle/gen_golden.py) — are the pin.
Tolerance: 2e-5 max-abs in fp32 (both sides are fp32 with different op orderings; the reference's
own fp32-vs-fp64 gap on these cases is ~1e-6) and 5e-6 against the oracle run in fp64.
"""
import numpy as np
import pytest
import torch

from oracle import smx_oracle as O
from tests import _golden as G


def run_oracle(fx, dtype):
    c = fx.cfg
    x = fx.x.to(dtype)
    kind = c["kind"]
    if kind == "cell":
        return O.summary_mixing(x, fx.sd, "", mode=c["mode"], act=c["act"], use_layernorm=c["use_layernorm"],
                                src_padding_mask=fx.mask, sum_mask=fx.sum_mask)
    if kind == "vanilla":
        return O.vanilla_nn(x, fx.sd, "", c["act"])
    if kind == "conv_module":
        return O.convolution_module(x, fx.sd, "", act=c["act"], mask=fx.mask.unsqueeze(-1), causal=c["causal"],
                                    masked_false_or_true=False, chunk_size=c["chunk_size"])
    if kind == "conformer_layer":
        return O.conformer_layer(x, fx.sd, "", act=c["act"], mode=c["mode"], use_layernorm=c["use_layernorm"],
                                 src_key_padding_mask=fx.mask)
    if kind == "conformer_encoder":
        return O.conformer_encoder(x, fx.sd, c["num_layers"], act=c["act"], mode=c["mode"],
                                   use_layernorm=c["use_layernorm"], src_mask=fx.sum_mask,
                                   src_key_padding_mask=fx.mask, chunk_size=c["chunk_size"])
    if kind == "branchformer_encoder":
        return O.branchformer_encoder(x, fx.sd, c["num_layers"], act=c["act"], gate_act=c["gate_act"], mode=c["mode"],
                                      src_key_padding_mask=fx.mask)
    raise AssertionError(kind)


@pytest.mark.parametrize("name", G.names())
def test_oracle_matches_reference_golden(name):
    fx = G.Fixture(name)
    y32 = run_oracle(fx, torch.float32)
    y64 = run_oracle(fx, torch.float64)
    assert y32.shape == fx.y.shape
    assert float((y32 - fx.y).abs().max()) < 2e-5
    assert float((y64 - fx.y.double()).abs().max()) < 5e-6


def test_oracle_masks_match_reference():
    z = np.load(G.GOLDEN_DIR + "/masks.npz")
    pad = O.padding_mask_from_wav_len(torch.from_numpy(z["wav_len"]), 37)
    assert np.array_equal(pad.numpy(), z["padding_mask"])
    for key in z.files:
        if key.startswith("chunk_"):
            _, cs, lc = key.split("_")
            m = O.chunk_mask(37, int(cs), None if lc == "None" else int(lc))
            assert np.array_equal(m.numpy(), z[key]), key


def test_padded_frames_constant_and_nonzero():
    """padded-frame outputs are non-zero and identical across an utterance's padded frames."""
    fx = G.Fixture("cell_sm_h4_swish")
    y = fx.y
    assert float(y[3, 7:].abs().max()) > 0
    assert float((y[3, 7:] - y[3, 7:8]).abs().max()) < 1e-6


def test_combiner_split_is_exact():
    """act(W_c [local; mu] + b) == act(W_cl local + (W_cs mu + b)) — the algebraic split the kernels use."""
    fx = G.Fixture("cell_sm_h4_swish")
    sd = {k: v.double() for k, v in fx.sd.items()}
    x = fx.x.double()
    m = fx.mask.double().unsqueeze(-1)
    local = O.layer_norm(O.vanilla_nn(x, sd, "local_proj.", "swish") * m, sd["local_norm.weight"], sd["local_norm.bias"])
    s = O.vanilla_nn(x, sd, "summary_proj.", "swish") * m
    mu = O.layer_norm(s.sum(1) / m.sum(1), sd["summary_norm.weight"], sd["summary_norm.bias"])
    W, b = sd["summary_local_merging.linear.w.weight"], sd["summary_local_merging.linear.w.bias"]
    y = O.activation("swish", local @ W[:, :64].T + (mu @ W[:, 64:].T + b).unsqueeze(1))
    assert float((y - fx.y.double()).abs().max()) < 5e-6


def test_mhsa_comparison_arm_oracle_matches_reference():
    """oracle.conformer_encoder_mhsa (the self-attention comparison arm of BASELINE.json configs[4], tools/rtf_sweep.py) against
    the unmodified reference's ConformerEncoder(attention_type='regularMHA') output (fixture by oracle/gen_golden_tile.py --mhsa).
    The state_dict is rebuilt from the seeded stream with the reference's key names and shapes."""
    import json
    import os

    import numpy as np

    from oracle.seeded import seeded_input, seeded_param

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mhsa", "mhsa_conformer_enc.npz"))
    cfg = json.loads(bytes(z["cfg"]).decode())
    D, F, k = cfg["d_model"], cfg["d_ffn"], cfg["kernel_size"]
    shapes = {"in_proj_weight": (3 * D, D), "in_proj_bias": (3 * D,), "out_proj.weight": (D, D), "out_proj.bias": (D,),
              "bottleneck.0.weight": (2 * D, D, 1), "bottleneck.0.bias": (2 * D,), "conv.weight": (D, 1, k), "conv.bias": (D,),
              "after_conv.2.weight": (D, D), "ffn.0.weight": (F, D), "ffn.0.bias": (F,), "ffn.3.weight": (D, F)}
    sd = {}
    for key in [str(s) for s in z["keys"]]:
        shape = next((v for s, v in shapes.items() if key.endswith(s)), (D,))
        sd[key] = seeded_param(cfg["seed_w"], key, shape)
    x = seeded_input(cfg["seed_x"], cfg["B"], cfg["T"], D)
    y = O.conformer_encoder_mhsa(x, sd, cfg["num_layers"], cfg["nhead"], act=cfg["act"], key_padding_mask=torch.from_numpy(z["pad"]))
    assert float((y - torch.from_numpy(z["y32"])).abs().max()) < 2e-5
