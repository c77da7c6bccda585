"""Golden vectors at TILE-ALIGNED model sizes (D = 256, 512), generated from the UNMODIFIED reference.

    python oracle/gen_golden_tile.py        # writes tests/golden/tile/*.npz   (build container only)

tests/golden/*.npz (gen_golden.py) are D = 64 toys that never reach the fused tcgen05 kernels' shapes; these fixtures do.
Weights and inputs are seed-defined (oracle/seeded.py), so a fixture holds only the config, the seeds, the masks and
  y32   the reference module's fp32 output on the (bf16-representable) input                   -- the parity target
  y16   the SAME reference module cast to bfloat16 (.bfloat16(), bf16 input), output as float  -- the reference's own
        bf16-vs-fp32 error on these inputs is |y16 - y32|: the bound SURVEY.md 8d sanctions for a bf16 arm.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "sbshim"))
sys.path.insert(0, os.path.dirname(HERE))

from speechbrain.nnet.activations import Swish  # noqa: E402
from speechbrain.nnet.summary_mixing import SummaryMixing  # noqa: E402
from speechbrain.lobes.models.transformer.Conformer import ConformerEncoder, ConformerEncoderLayer, ConvolutionModule  # noqa: E402
from speechbrain.lobes.models.transformer.Branchformer import BranchformerEncoder  # noqa: E402

from oracle.seeded import fill_module, seeded_input  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "tile")
ACTS = {"swish": Swish, "gelu": nn.GELU}


def lens_mask(T, lens):
    return torch.arange(T)[None, :] < torch.tensor(lens)[:, None]


def emit(name, cfg, module, x, mask, run):
    module.eval()
    with torch.no_grad():
        y32 = run(module.float(), x.float()).float().contiguous()
        y16 = run(module.bfloat16(), x.to(torch.bfloat16)).float().contiguous()
        module.float()
    cfg = dict(cfg, reference_commit="d1b1f42", torch=torch.__version__, B=x.shape[0], T=x.shape[1])
    e16 = float((y16 - y32).abs().max())
    r16 = float((y16 - y32).norm() / y32.norm())
    cfg["ref_bf16_maxabs"], cfg["ref_bf16_rel_l2"], cfg["y_absmax"] = e16, r16, float(y32.abs().max())
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), y32=y32.numpy(),
                        y16=y16.to(torch.bfloat16).view(torch.int16).numpy(),  # bf16 bit patterns
                        mask=mask.numpy(), cfg=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8))
    print(f"{name}: y {tuple(y32.shape)} |y|max {cfg['y_absmax']:.3f}; reference bf16-vs-fp32 max-abs {e16:.3e} rel-L2 {r16:.3e}")


def main():
    # ---- cfg2 dims (BASELINE.json configs[1]): D=256, h=4, hidden 256, d_ffn=1024, Swish -----------------------
    D, h, F = 256, 4, 1024
    B, T = 2, 200                                 # two tiles per utterance, ragged last tile, one padded utterance
    mask = lens_mask(T, [200, 131])
    x = seeded_input(11, B, T, D)
    for act in ("swish", "gelu"):
        for nhead in ((4, 1) if act == "swish" else (4,)):
            cell = SummaryMixing(D, nhead, [D], D, [D], D, activation=ACTS[act], mode="SummaryMixing")
            fill_module(cell, 21)
            cfg = dict(kind="cell", enc_dim=D, nhead=nhead, local_proj_hid_dim=[D], local_proj_out_dim=D, summary_hid_dim=[D],
                       summary_out_dim=D, act=act, mode="SummaryMixing", use_layernorm=True, seed_w=21, seed_x=11)
            emit(f"cell_d256_h{nhead}_{act}", cfg, cell, x, mask, lambda m, xx: m(xx, src_padding_mask=mask))
    for mode in ("SummaryMixing-lite", "SummaryMixing-fast"):
        cell = SummaryMixing(D, 4, [D], D, [D], D, activation=Swish, mode=mode, use_layernorm=(mode != "SummaryMixing-fast"))
        fill_module(cell, 22)
        cfg = dict(kind="cell", enc_dim=D, nhead=4, local_proj_hid_dim=[D], local_proj_out_dim=D, summary_hid_dim=[D],
                   summary_out_dim=D, act="swish", mode=mode, use_layernorm=(mode != "SummaryMixing-fast"), seed_w=22, seed_x=11)
        tag = mode.replace("SummaryMixing-", "")
        emit(f"cell_d256_h4_{tag}", cfg, cell, x, mask, lambda m, xx: m(xx, src_padding_mask=mask).contiguous())
    cm = ConvolutionModule(D, 31, True, Swish, 0.0, masked_false_or_true=False)
    fill_module(cm, 23)
    emit("convmod_d256", dict(kind="conv_module", input_size=D, kernel_size=31, act="swish", seed_w=23, seed_x=11), cm, x, mask,
         lambda m, xx: m(xx, mask.unsqueeze(-1)))
    layer = ConformerEncoderLayer(D, F, h, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                                  summary_hid_dim=[D], mode="SummaryMixing")
    fill_module(layer, 24)
    cfgl = dict(kind="conformer_layer", d_model=D, d_ffn=F, nhead=h, kernel_size=31, act="swish", mode="SummaryMixing",
                local_proj_hid_dim=[D], local_proj_out_dim=D, summary_hid_dim=[D], use_layernorm=True, seed_w=24, seed_x=11)
    emit("conformer_layer_d256", cfgl, layer, x, mask, lambda m, xx: m(xx, src_key_padding_mask=mask)[0])
    enc = ConformerEncoder(3, D, F, h, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                           summary_hid_dim=[D], mode="SummaryMixing")
    fill_module(enc, 25)
    cfge = dict(cfgl, kind="conformer_encoder", num_layers=3, seed_w=25)
    emit("conformer_enc_d256_3l", cfge, enc, x, mask, lambda m, xx: m(xx, src_key_padding_mask=mask)[0])

    # ---- cfg3 dims (conformer_summarymixing.yaml:113-125): D=512, h=8, d_ffn=2048, hidden/out 512 ---------------
    D, h, F = 512, 8, 2048
    B, T = 2, 136
    mask = lens_mask(T, [136, 77])
    x = seeded_input(12, B, T, D)
    layer = ConformerEncoderLayer(D, F, h, 31, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                                  summary_hid_dim=[D], mode="SummaryMixing")
    fill_module(layer, 31)
    cfgl = dict(kind="conformer_layer", d_model=D, d_ffn=F, nhead=h, kernel_size=31, act="swish", mode="SummaryMixing",
                local_proj_hid_dim=[D], local_proj_out_dim=D, summary_hid_dim=[D], use_layernorm=True, seed_w=31, seed_x=12)
    emit("conformer_layer_d512", cfgl, layer, x, mask, lambda m, xx: m(xx, src_key_padding_mask=mask)[0])

    # ---- cfg4 dims (branchformer_summarymixing.yaml:112-127, mode lite): D=512, h=1, csgu 3072, k=31 ------------
    enc = BranchformerEncoder(2, D, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[D], local_proj_out_dim=D,
                              summary_hid_dim=[D], summary_out_dim=D, mode="SummaryMixing-lite")
    fill_module(enc, 41)
    cfgb = dict(kind="branchformer_encoder", num_layers=2, d_model=D, nhead=1, kernel_size=31, csgu_linear_units=3072, act="gelu",
                gate_act="identity", mode="SummaryMixing-lite", local_proj_hid_dim=[D], local_proj_out_dim=D, summary_hid_dim=[D],
                summary_out_dim=D, seed_w=41, seed_x=12)
    emit("branchformer_enc_d512_lite_2l", cfgb, enc, x, mask, lambda m, xx: m(xx, src_key_padding_mask=mask)[0])
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"total {total / 1e6:.2f} MB in {OUT}")


if __name__ == "__main__" and "--mhsa" not in sys.argv:
    main()


def mhsa_fixture():
    """The self-attention Conformer (attention_type='regularMHA', Conformer.py:425-429): the comparison arm of BASELINE.json
    configs[4].  Pins oracle.conformer_encoder_mhsa; D = 64 toy (the arm is a torch-library baseline, not a product kernel)."""
    torch.manual_seed(77)
    enc = ConformerEncoder(2, 64, 128, 4, 31, attention_type="regularMHA").eval()
    fill_module(enc, 78)
    x = seeded_input(79, 3, 70, 64)
    pad = ~lens_mask(70, [70, 41, 12])
    with torch.no_grad():
        y = enc(x, src_key_padding_mask=pad)[0]
    cfg = dict(kind="conformer_encoder_mhsa", num_layers=2, d_model=64, d_ffn=128, nhead=4, kernel_size=31, act="swish", seed_w=78, seed_x=79,
               B=3, T=70, reference_commit="d1b1f42", torch=torch.__version__)
    np.savez_compressed(os.path.join(os.path.dirname(OUT), "mhsa", "mhsa_conformer_enc.npz"), y32=y.numpy(), pad=pad.numpy(),
                        keys=np.array(sorted(enc.state_dict().keys())), cfg=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8))
    print("mhsa_conformer_enc: ok", tuple(y.shape))


if __name__ == "__main__" and "--mhsa" in sys.argv:
    mhsa_fixture()
