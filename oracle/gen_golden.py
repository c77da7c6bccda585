"""Generate golden input/output vectors from the UNMODIFIED reference (run in the build container).

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference (SamsungLabs/SummaryMixing @ d1b1f42) is imported from /root/reference through the
SpeechBrain stand-in ``oracle/sbshim`` — no reference source is copied.  Each fixture holds: the
config, the seeded input, masks, the reference module's ``state_dict`` (after a seeded perturbation so
that LayerNorm gains/biases and zero-initialised biases are exercised) and the reference output in
fp32 (the reference's own arithmetic) — plus an fp64 run of the same module for tolerance headroom.
Fixtures are what ``tests/`` on the GPU box compare against (/root/reference does not travel).
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

from speechbrain.nnet.activations import Swish  # noqa: E402
from speechbrain.nnet.summary_mixing import SummaryMixing  # noqa: E402
from speechbrain.lobes.models.VanillaNN import VanillaNN  # noqa: E402
from speechbrain.lobes.models.transformer.Conformer import (  # noqa: E402
    ConformerEncoder,
    ConformerEncoderLayer,
    ConvolutionModule,
)
from speechbrain.lobes.models.transformer.Branchformer import BranchformerEncoder  # noqa: E402
from speechbrain.lobes.models.transformer.TransformerASR import (  # noqa: E402
    make_transformer_src_mask,
    make_transformer_src_tgt_masks,
)
from speechbrain.utils.dynamic_chunk_training import DynChunkTrainConfig  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
ACTS = {"swish": Swish, "gelu": nn.GELU, "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU}
REF_COMMIT = "d1b1f42"
SHARED_INPUTS = {}


def perturb(module: nn.Module, seed: int) -> None:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("decay_constant"):
                continue
            if p.dim() >= 2:
                p.add_(0.02 * torch.randn(p.shape, generator=g))
            else:
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def prefix_mask(B, T, lens):
    return torch.arange(T)[None, :] < torch.tensor(lens)[:, None]


def save(name, cfg, module, inputs: dict, run):
    """run(module, dtype) -> output tensor.  Stores the fp32 output (the reference's own arithmetic)
    and records how far the reference's fp64 run is from it (cfg.ref_fp32_vs_fp64_maxabs)."""
    module.eval()
    with torch.no_grad():
        y32 = run(module.float(), torch.float32).contiguous()
        try:
            y64 = run(module.double(), torch.float64).contiguous()
        except RuntimeError:  # the reference casts sum_mask with .float() (summary_mixing.py:189): no fp64 run
            y64 = None
        module.float()
    arrays = {"y": y32.numpy()}
    err64 = None if y64 is None else float((y32.double() - y64).abs().max())
    for k, v in inputs.items():
        if k == "x":  # inputs are shared between fixtures: stored once in _inputs.npz under v[0]
            SHARED_INPUTS[v[0]] = v[1].numpy()
            arrays["in.x_ref"] = np.frombuffer(v[0].encode(), dtype=np.uint8)
        else:
            arrays["in." + k] = v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    for k, v in module.state_dict().items():
        arrays["sd." + k] = v.detach().float().numpy()
    cfg = dict(cfg, reference_commit=REF_COMMIT, torch=torch.__version__, ref_fp32_vs_fp64_maxabs=err64)
    arrays["cfg"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    os.makedirs(OUT, exist_ok=True)
    np.savez(os.path.join(OUT, name + ".npz"), **arrays)
    print(f"""{name}: y {tuple(y32.shape)} |y|max {y32.abs().max():.3f} fp32-vs-fp64 """
          f"""{'n/a' if err64 is None else format(err64, '.2e')}""")


def cell_cases():
    # BASELINE.json configs[0]: (B=4,T=128,D=64), lens per SURVEY.md 8d.1
    B, T, D = 4, 128, 64
    x1 = torch.randn(B, T, D, generator=torch.Generator().manual_seed(0))
    mask1 = prefix_mask(B, T, [128, 100, 64, 7])
    x2 = torch.randn(3, 37, D, generator=torch.Generator().manual_seed(9))
    mask2 = prefix_mask(3, 37, [37, 20, 1])
    cfg1_set = {("SummaryMixing", 1, "gelu"), ("SummaryMixing", 4, "swish"), ("SummaryMixing-lite", 4, "gelu"),
                ("SummaryMixing-fast", 4, "swish"), ("SummaryMixing-fast", 1, "gelu"), ("SummaryMixing-expdecay", 4, "gelu")}
    seed = 100
    for mode in ("SummaryMixing", "SummaryMixing-lite", "SummaryMixing-fast", "SummaryMixing-expdecay"):
        for nhead in (1, 4):
            for act in ("gelu", "swish"):
                if act == "swish" and mode in ("SummaryMixing-lite", "SummaryMixing-expdecay"):
                    continue
                seed += 1
                if (mode, nhead, act) in cfg1_set:
                    xname, x, mask = "x_cfg1", x1, mask1
                else:
                    xname, x, mask = "x_small", x2, mask2
                torch.manual_seed(seed)
                sm = SummaryMixing(D, nhead, [64], 64, [64], 64, activation=ACTS[act], mode=mode)
                perturb(sm, seed)
                cfg = dict(kind="cell", enc_dim=D, nhead=nhead, local_proj_hid_dim=[64], local_proj_out_dim=64,
                           summary_hid_dim=[64], summary_out_dim=64, act=act, mode=mode, use_layernorm=True)
                tag = mode.replace("SummaryMixing", "sm").replace("-", "_")
                save(f"cell_{tag}_h{nhead}_{act}", cfg, sm, {"x": (xname, x), "mask": mask},
                     lambda m, dt: m(x.to(dt), src_padding_mask=mask))
    # use_layernorm=False (transducer recipes), unequal dims, two hidden layers, no mask
    torch.manual_seed(201)
    sm = SummaryMixing(D, 4, [32, 48], 32, [96], 64, activation=ACTS["relu"], mode="SummaryMixing", use_layernorm=False)
    perturb(sm, 201)
    cfg = dict(kind="cell", enc_dim=D, nhead=4, local_proj_hid_dim=[32, 48], local_proj_out_dim=32, summary_hid_dim=[96],
               summary_out_dim=64, act="relu", mode="SummaryMixing", use_layernorm=False)
    x = x1
    save("cell_sm_h4_relu_noln_deep_nomask", cfg, sm, {"x": ("x_cfg1", x)}, lambda m, dt: m(x.to(dt)))
    # the reference unit test's own shapes (tests/unittests/test_summary_mixing.py:5-57): rand(8,10,64) seed 666
    torch.manual_seed(666)
    xt = torch.rand(8, 10, 64)
    for mode in ("SummaryMixing", "SummaryMixing-lite"):
        for nhead in (1, 4):
            sm = SummaryMixing(enc_dim=64, nhead=nhead, local_proj_hid_dim=[32], local_proj_out_dim=32,
                               summary_out_dim=64, mode=mode)
            perturb(sm, 300 + nhead)
            cfg = dict(kind="cell", enc_dim=64, nhead=nhead, local_proj_hid_dim=[32], local_proj_out_dim=32,
                       summary_hid_dim=[512], summary_out_dim=64, act="gelu", mode=mode, use_layernorm=True)
            tag = mode.replace("SummaryMixing", "sm").replace("-", "_")
            save(f"cell_reftest_{tag}_h{nhead}", cfg, sm, {"x": ("x_reftest", xt)}, lambda m, dt: m(xt.to(dt)))
    # sum_mask (dynamic-chunk) path, full + fast, summary_mixing.py:235-246,292-294
    Ts = 50
    xs = torch.randn(3, Ts, D, generator=torch.Generator().manual_seed(5))
    masks = prefix_mask(3, Ts, [50, 33, 9])
    for mode, lc in (("SummaryMixing", None), ("SummaryMixing-fast", 1)):
        dc = DynChunkTrainConfig(8, lc)
        smask = make_transformer_src_mask(xs, False, False, dc)
        torch.manual_seed(400)
        sm = SummaryMixing(D, 4, [64], 64, [64], 64, activation=Swish, mode=mode, use_layernorm=(mode != "SummaryMixing-fast"))
        perturb(sm, 400)
        cfg = dict(kind="cell", enc_dim=D, nhead=4, local_proj_hid_dim=[64], local_proj_out_dim=64, summary_hid_dim=[64],
                   summary_out_dim=64, act="swish", mode=mode, use_layernorm=(mode != "SummaryMixing-fast"),
                   chunk_size=8, left_context_size=lc)
        tag = mode.replace("SummaryMixing", "sm").replace("-", "_")
        save(f"cell_{tag}_h4_swish_summask", cfg, sm, {"x": ("x_summask", xs), "mask": masks, "sum_mask": smask},
             lambda m, dt: m(xs.to(dt), sum_mask=smask, src_padding_mask=masks))


def vanilla_cases():
    x = torch.randn(2, 9, 48, generator=torch.Generator().manual_seed(7))
    for n_split in (1, 3):
        torch.manual_seed(500 + n_split)
        nn_ = VanillaNN(input_shape=[None, None, 48], activation=nn.LeakyReLU, dnn_blocks=3, dnn_neurons=[24, 36, 12], n_split=n_split)
        perturb(nn_, 500 + n_split)
        cfg = dict(kind="vanilla", act="leaky_relu", n_split=n_split, dnn_neurons=[24, 36, 12], input_size=48)
        save(f"vanilla_split{n_split}", cfg, nn_, {"x": ("x_vanilla", x)}, lambda m, dt: m(x.to(dt)))


def conformer_cases():
    B, T, D = 3, 70, 64
    x = torch.randn(B, T, D, generator=torch.Generator().manual_seed(1))
    mask = prefix_mask(B, T, [70, 41, 12])
    # convolution module alone (+mask, SummaryMixing convention), causal, and dynamic-chunk variants
    for tag, causal, chunk in (("plain", False, None), ("causal", True, None), ("dcconv", False, 16)):
        torch.manual_seed(600)
        cm = ConvolutionModule(D, 31, True, Swish, 0.0, causal=causal, masked_false_or_true=False)
        perturb(cm, 600)
        cfg = dict(kind="conv_module", input_size=D, kernel_size=31, act="swish", causal=causal, chunk_size=chunk)
        dc = None if chunk is None else DynChunkTrainConfig(chunk, None)
        save(f"convmod_{tag}", cfg, cm, {"x": ("x_conformer", x), "mask": mask},
             lambda m, dt: m(x.to(dt), mask.unsqueeze(-1), dynchunktrain_config=dc))
    # one layer
    torch.manual_seed(610)
    layer = ConformerEncoderLayer(D, 128, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[64],
                                  local_proj_out_dim=64, summary_hid_dim=[64], mode="SummaryMixing")
    perturb(layer, 610)
    cfg = dict(kind="conformer_layer", d_model=D, d_ffn=128, nhead=4, kernel_size=31, act="swish", mode="SummaryMixing",
               local_proj_hid_dim=[64], local_proj_out_dim=64, summary_hid_dim=[64], use_layernorm=True)
    save("conformer_layer", cfg, layer, {"x": ("x_conformer", x), "mask": mask}, lambda m, dt: m(x.to(dt), src_key_padding_mask=mask)[0])
    # encoders: full h=4 swish; full h=1 gelu; fast no-LN with dynchunk (transducer-recipe shape)
    for tag, nhead, act, mode, uln, chunk in (
        ("sm_h4", 4, "swish", "SummaryMixing", True, None),
        ("sm_h1_gelu", 1, "gelu", "SummaryMixing", True, None),
        ("lite_h4", 4, "swish", "SummaryMixing-lite", True, None),
        ("fast_noln_dynchunk", 4, "swish", "SummaryMixing-fast", False, 16),
    ):
        torch.manual_seed(620)
        enc = ConformerEncoder(2, D, 128, nhead, 31, activation=ACTS[act], attention_type="SummaryMixing",
                               local_proj_hid_dim=[64], local_proj_out_dim=64, summary_hid_dim=[64], mode=mode,
                               use_layernorm=uln)
        perturb(enc, 620)
        cfg = dict(kind="conformer_encoder", num_layers=2, d_model=D, d_ffn=128, nhead=nhead, kernel_size=31, act=act,
                   mode=mode, local_proj_hid_dim=[64], local_proj_out_dim=64, summary_hid_dim=[64], use_layernorm=uln,
                   chunk_size=chunk)
        dc = None if chunk is None else DynChunkTrainConfig(chunk, None)
        smask = None if dc is None else make_transformer_src_mask(x, False, False, dc)
        ins = {"x": ("x_conformer", x), "mask": mask}
        if smask is not None:
            ins["sum_mask"] = smask
        save(f"conformer_enc_{tag}", cfg, enc, ins,
             lambda m, dt: m(x.to(dt), src_mask=smask, src_key_padding_mask=mask, dynchunktrain_config=dc)[0])


def branchformer_cases():
    B, T, D = 3, 70, 64
    x = torch.randn(B, T, D, generator=torch.Generator().manual_seed(2))
    mask = prefix_mask(B, T, [70, 41, 12])
    for tag, mode in (("lite", "SummaryMixing-lite"), ("full", "SummaryMixing")):
        torch.manual_seed(700)
        enc = BranchformerEncoder(2, D, 1, 31, csgu_linear_units=192, local_proj_hid_dim=[64], local_proj_out_dim=64,
                                  summary_hid_dim=[64], summary_out_dim=64, mode=mode)
        perturb(enc, 700)
        with torch.no_grad():  # CSGU conv weights are ~1e-6 at init; make the conv matter
            for n, p in enc.named_parameters():
                if "csgu.conv.conv.weight" in n:
                    p.add_(0.2 * torch.randn(p.shape, generator=torch.Generator().manual_seed(701)))
        cfg = dict(kind="branchformer_encoder", num_layers=2, d_model=D, nhead=1, kernel_size=31, csgu_linear_units=192,
                   act="gelu", gate_act="identity", mode=mode, local_proj_hid_dim=[64], local_proj_out_dim=64,
                   summary_hid_dim=[64], summary_out_dim=64)
        save(f"branchformer_enc_{tag}", cfg, enc, {"x": ("x_branchformer", x), "mask": mask},
             lambda m, dt: m(x.to(dt), src_key_padding_mask=mask)[0])


def mask_cases():
    src = torch.zeros(4, 37, 8)
    wav_len = torch.tensor([1.0, 0.73, 0.5, 0.051])
    pad, _, _, _ = make_transformer_src_tgt_masks(src, None, wav_len, masked_false_or_true=False)
    arrays = {"wav_len": wav_len.numpy(), "padding_mask": pad.numpy()}
    for cs, lc in ((8, None), (8, 1), (5, 0), (16, 2)):
        m = make_transformer_src_mask(src, False, False, DynChunkTrainConfig(cs, lc))
        arrays[f"chunk_{cs}_{lc}"] = m.numpy()
    np.savez(os.path.join(OUT, "masks.npz"), **arrays)
    print("masks: ok")


if __name__ == "__main__":
    cell_cases()
    vanilla_cases()
    conformer_cases()
    branchformer_cases()
    mask_cases()
    np.savez(os.path.join(OUT, "_inputs.npz"), **SHARED_INPUTS)
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"total {total / 1e6:.2f} MB in {OUT}")
