"""ctypes binding of libsmx.so (the C ABI declared in include/smx.h).

The product path is CUDA only: if the library is missing, or no sm_100 device is present, calls fail
loudly — there is no CPU fallback (the CPU oracle under oracle/ is test infrastructure and is never
imported from here).
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libsmx.so")

SMX_MAX_BLOCKS = 4
F32, BF16 = 0, 1
ACT_IDENTITY, ACT_SWISH, ACT_GELU, ACT_RELU, ACT_LEAKY_RELU, ACT_TANH, ACT_SIGMOID, ACT_GELU_TANH = range(8)
MODE_FULL, MODE_LITE, MODE_FAST, MODE_EXPDECAY = range(4)
MODES = {
    "SummaryMixing": MODE_FULL,
    "SummaryMixing-lite": MODE_LITE,
    "SummaryMixing-fast": MODE_FAST,
    "SummaryMixing-expdecay": MODE_EXPDECAY,
}
STATUS = {0: "SMX_OK", -1: "SMX_ERR_BAD_ARG", -2: "SMX_ERR_UNSUPPORTED", -3: "SMX_ERR_ALIGNMENT",
          -4: "SMX_ERR_WORKSPACE", -5: "SMX_ERR_CUDA", -6: "SMX_ERR_ARCH"}

fp = C.c_void_p  # const float* (device)


class Linear(C.Structure):
    _fields_ = [("w", fp), ("b", fp), ("in_dim", C.c_int32), ("out_dim", C.c_int32), ("n_split", C.c_int32),
                ("_pad", C.c_int32)]


class CellWeights(C.Structure):
    _fields_ = [("mode", C.c_int32), ("act", C.c_int32), ("use_layernorm", C.c_int32), ("enc_dim", C.c_int32),
                ("local_out_dim", C.c_int32), ("summary_out_dim", C.c_int32), ("n_local", C.c_int32),
                ("n_summary", C.c_int32), ("local", Linear * SMX_MAX_BLOCKS), ("summary", Linear * SMX_MAX_BLOCKS),
                ("global_proj", Linear), ("merge", Linear), ("local_norm_w", fp), ("local_norm_b", fp),
                ("summary_norm_w", fp), ("summary_norm_b", fp), ("packed", fp), ("decay_constant", C.c_float),
                ("_pad", C.c_int32), ("prenorm_w", fp), ("prenorm_b", fp)]


class LinearGrad(C.Structure):
    _fields_ = [("dw", fp), ("db", fp)]


class CellGrads(C.Structure):
    _fields_ = [("local", LinearGrad * SMX_MAX_BLOCKS), ("summary", LinearGrad * SMX_MAX_BLOCKS), ("merge", LinearGrad),
                ("local_norm_dw", fp), ("local_norm_db", fp), ("summary_norm_dw", fp), ("summary_norm_db", fp),
                ("global_proj", LinearGrad)]


class FFNGrads(C.Structure):
    _fields_ = [("ln_dw", fp), ("ln_db", fp), ("w1", LinearGrad), ("w2", LinearGrad), ("out_ln_dw", fp), ("out_ln_db", fp)]


class Dropout(C.Structure):
    """smx_dropout: p and the seed of the counter-based masks (include/smx.h)."""
    _fields_ = [("p", C.c_float), ("seed", C.c_uint64)]


class ConvModGrads(C.Structure):
    _fields_ = [("ln_dw", fp), ("ln_db", fp), ("bottleneck", LinearGrad), ("dw_dw", fp), ("dw_db", fp), ("after_ln_dw", fp),
                ("after_ln_db", fp), ("out", LinearGrad)]


class ConvBranchGrads(C.Structure):
    _fields_ = [("pre", LinearGrad), ("post", LinearGrad), ("csgu_ln_dw", fp), ("csgu_ln_db", fp), ("csgu_dw_dw", fp),
                ("csgu_dw_db", fp), ("csgu_linear", LinearGrad)]


class FFNWeights(C.Structure):
    _fields_ = [("ln_w", fp), ("ln_b", fp), ("w1", Linear), ("w2", Linear), ("packed", fp)]


class ConvModWeights(C.Structure):
    _fields_ = [("ln_w", fp), ("ln_b", fp), ("bottleneck", Linear), ("dw_w", fp), ("dw_b", fp), ("after_ln_w", fp),
                ("after_ln_b", fp), ("out", Linear), ("packed", fp), ("kernel_size", C.c_int32), ("causal", C.c_int32)]


class ConformerLayerWeights(C.Structure):
    _fields_ = [("ffn1", FFNWeights), ("ffn2", FFNWeights), ("norm1_w", fp), ("norm1_b", fp), ("norm2_w", fp),
                ("norm2_b", fp), ("cell", CellWeights), ("conv", ConvModWeights), ("act", C.c_int32),
                ("_pad", C.c_int32)]


class ConvBranchWeights(C.Structure):
    _fields_ = [("pre", Linear), ("post", Linear), ("csgu_ln_w", fp), ("csgu_ln_b", fp), ("csgu_dw_w", fp),
                ("csgu_dw_b", fp), ("csgu_linear", Linear), ("kernel_size", C.c_int32), ("act", C.c_int32),
                ("gate_act", C.c_int32), ("_pad", C.c_int32)]


class BranchformerLayerWeights(C.Structure):
    _fields_ = [("norm_mhsa_w", fp), ("norm_mhsa_b", fp), ("norm_conv_w", fp), ("norm_conv_b", fp),
                ("cell", CellWeights), ("branch", ConvBranchWeights), ("n_merge", C.c_int32), ("act", C.c_int32),
                ("merge", Linear * SMX_MAX_BLOCKS), ("packed", C.c_void_p)]


class FbankDesc(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("n_mels", C.c_int32), ("win_length_ms", C.c_float),
                ("hop_length_ms", C.c_float), ("f_min", C.c_float), ("f_max", C.c_float), ("amin", C.c_float), ("top_db", C.c_float)]


class SmxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsmx: {STATUS.get(code, code)}: {msg}")
        self.code = code


_lib = None

_i, _i64, _sz, _vp, _f = C.c_int, C.c_int64, C.c_size_t, C.c_void_p, C.c_float
_PROTOS = {
    "smx_version": (C.c_int, []),
    "smx_last_error": (C.c_char_p, []),
    "smx_launch_count": (C.c_uint64, []),
    "smx_tc_launch_count": (C.c_uint64, []),
    "smx_struct_size": (_sz, [_i]),
    "smx_cell_packed_bytes": (_sz, [C.POINTER(CellWeights)]),
    "smx_cell_pack": (_i, [C.POINTER(CellWeights), _vp, _sz, _vp]),
    "smx_cell_pack_prenorm": (_i, [C.POINTER(CellWeights), fp, fp, _vp]),
    "smx_ffn_packed_bytes": (_sz, [C.POINTER(FFNWeights)]),
    "smx_ffn_pack": (_i, [C.POINTER(FFNWeights), _vp, _sz, _vp]),
    "smx_branchformer_packed_bytes": (_sz, [C.POINTER(BranchformerLayerWeights)]),
    "smx_branchformer_pack": (_i, [C.POINTER(BranchformerLayerWeights), _vp, _sz, _vp]),
    "smx_convmod_packed_bytes": (_sz, [C.POINTER(ConvModWeights)]),
    "smx_convmod_pack": (_i, [C.POINTER(ConvModWeights), _vp, _sz, _vp]),
    "smx_layernorm_fwd": (_i, [_i, _i64, _i, _vp, _vp, _vp, _f, _vp, _vp]),
    "smx_vanilla_nn_workspace_bytes": (_sz, [C.POINTER(Linear), _i, _i, _i64]),
    "smx_vanilla_nn_fwd": (_i, [C.POINTER(Linear), _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "smx_summary_mixing_workspace_bytes": (_sz, [C.POINTER(CellWeights), _i, _i, _i, _i]),
    "smx_summary_mixing_fwd": (_i, [C.POINTER(CellWeights), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "smx_summary_mixing_bwd_workspace_bytes": (_sz, [C.POINTER(CellWeights), _i, _i, _i]),
    "smx_summary_mixing_bwd": (_i, [C.POINTER(CellWeights), _i, _i, _i, _vp, _vp, _vp, _vp, C.POINTER(CellGrads), _vp, _sz,
                                    _vp]),
    "smx_vanilla_nn_bwd_workspace_bytes": (_sz, [C.POINTER(Linear), _i, _i, _i64]),
    "smx_vanilla_nn_bwd": (_i, [C.POINTER(Linear), _i, _i, _i, _i64, _vp, _vp, _vp, C.POINTER(LinearGrad), _vp, _sz, _vp]),
    "smx_layernorm_bwd_workspace_bytes": (_sz, [_i, _i64, _i]),
    "smx_layernorm_bwd": (_i, [_i, _i64, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "smx_ffn_bwd_workspace_bytes": (_sz, [C.POINTER(FFNWeights), _i, _i64, _i]),
    "smx_ffn_bwd": (_i, [C.POINTER(FFNWeights), _i, _i, _i64, _vp, _vp, _vp, _f, _vp, _vp, C.POINTER(FFNGrads), _vp, _sz,
                         _vp]),
    "smx_conv_module_bwd_workspace_bytes": (_sz, [C.POINTER(ConvModWeights), _i, _i, _i]),
    "smx_conv_module_bwd": (_i, [C.POINTER(ConvModWeights), _i, _i, _i, _i, _vp, _vp, _vp, _vp, C.POINTER(ConvModGrads),
                                 _vp, _sz, _vp]),
    "smx_ffn_train_workspace_bytes": (_sz, [C.POINTER(FFNWeights), _i, C.c_int64, _i]),
    "smx_ffn_train_fwd": (_i, [C.POINTER(FFNWeights), _i, _i, C.c_int64, _vp, _vp, _vp, C.c_float, C.POINTER(Dropout), _vp, _vp, _sz, _vp]),
    "smx_ffn_train_bwd": (_i, [C.POINTER(FFNWeights), _i, _i, C.c_int64, _vp, _vp, _vp, C.c_float, C.POINTER(Dropout), _vp, _vp,
                               C.POINTER(FFNGrads), _vp, _sz, _vp]),
    "smx_conv_module_train_workspace_bytes": (_sz, [C.POINTER(ConvModWeights), _i, _i, _i]),
    "smx_conv_module_train_fwd": (_i, [C.POINTER(ConvModWeights), _i, _i, _i, _i, _vp, _vp, C.POINTER(Dropout), _vp, _vp, _sz, _vp]),
    "smx_conv_module_train_bwd": (_i, [C.POINTER(ConvModWeights), _i, _i, _i, _i, _vp, _vp, C.POINTER(Dropout), _vp, _vp,
                                       C.POINTER(ConvModGrads), _vp, _sz, _vp]),
    "smx_summary_mixing_train_workspace_bytes": (_sz, [C.POINTER(CellWeights), _i, _i, _i]),
    "smx_summary_mixing_train_fwd": (_i, [C.POINTER(CellWeights), _i, _i, _i, _vp, _vp, C.POINTER(Dropout), _vp, _vp, _sz, _vp]),
    "smx_summary_mixing_train_bwd": (_i, [C.POINTER(CellWeights), _i, _i, _i, _vp, _vp, C.POINTER(Dropout), _vp, _vp,
                                          C.POINTER(CellGrads), _vp, _sz, _vp]),
    "smx_dropout_keep_mask": (_i, [C.POINTER(Dropout), _i, C.c_int64, _vp, _vp]),
    "smx_dropout_apply": (_i, [C.POINTER(Dropout), _i, _i, C.c_int64, _vp, _vp, _vp]),
    "smx_summary_mixing_masked_train_workspace_bytes": (_sz, [C.POINTER(CellWeights), _i, _i, _i]),
    "smx_summary_mixing_masked_train_fwd": (_i, [C.POINTER(CellWeights), _i, _i, _i, _vp, _vp, _vp, C.POINTER(Dropout), _vp, _vp, _sz, _vp]),
    "smx_summary_mixing_masked_train_bwd": (_i, [C.POINTER(CellWeights), _i, _i, _i, _vp, _vp, _vp, C.POINTER(Dropout), _vp, _vp,
                                                 C.POINTER(CellGrads), _vp, _sz, _vp]),
    "smx_conv_module_dcc_train_fwd": (_i, [C.POINTER(ConvModWeights), _i, _i, _i, _i, _i, _vp, _vp, C.POINTER(Dropout), _vp, _vp, _sz, _vp]),
    "smx_conv_module_dcc_train_bwd": (_i, [C.POINTER(ConvModWeights), _i, _i, _i, _i, _i, _vp, _vp, C.POINTER(Dropout), _vp, _vp,
                                           C.POINTER(ConvModGrads), _vp, _sz, _vp]),
    "smx_conv_branch_train_workspace_bytes": (_sz, [C.POINTER(ConvBranchWeights), _i, _i, _i]),
    "smx_conv_branch_train_fwd": (_i, [C.POINTER(ConvBranchWeights), _i, _i, _i, _vp, C.POINTER(Dropout), _vp, _vp, _sz, _vp]),
    "smx_conv_branch_train_bwd": (_i, [C.POINTER(ConvBranchWeights), _i, _i, _i, _vp, C.POINTER(Dropout), _vp, _vp,
                                       C.POINTER(ConvBranchGrads), _vp, _sz, _vp]),
    "smx_conv_module_workspace_bytes": (_sz, [C.POINTER(ConvModWeights), _i, _i, _i]),
    "smx_conv_module_fwd": (_i, [C.POINTER(ConvModWeights), _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "smx_ffn_workspace_bytes": (_sz, [C.POINTER(FFNWeights), _i, _i64]),
    "smx_ffn_fwd": (_i, [C.POINTER(FFNWeights), _i, _i, _i64, _vp, _vp, _vp, _f, _vp, _vp, _sz, _vp]),
    "smx_mixing_block_workspace_bytes": (_sz, [C.POINTER(CellWeights), _i, _i, _i, _i]),
    "smx_mixing_block_fwd": (_i, [C.POINTER(CellWeights), _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "smx_mixing_block_fwd_batch": (_i, [C.POINTER(CellWeights), _vp, _vp, _i, _i, _i, _i, C.POINTER(C.c_void_p), _vp,
                                        C.POINTER(C.c_void_p), _vp, _sz, _vp]),
    "smx_conformer_layer_workspace_bytes": (_sz, [C.POINTER(ConformerLayerWeights), _i, _i, _i, _i]),
    "smx_conformer_layer_fwd": (_i, [C.POINTER(ConformerLayerWeights), _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "smx_conformer_encoder_workspace_bytes": (_sz, [C.POINTER(ConformerLayerWeights), _i, _i, _i, _i, _i]),
    "smx_conformer_encoder_fwd": (_i, [C.POINTER(ConformerLayerWeights), _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp,
                                       _vp, C.POINTER(C.c_void_p), _vp, _sz, _vp]),
    "smx_branchformer_layer_workspace_bytes": (_sz, [C.POINTER(BranchformerLayerWeights), _i, _i, _i, _i]),
    "smx_branchformer_layer_fwd": (_i, [C.POINTER(BranchformerLayerWeights), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "smx_branchformer_encoder_workspace_bytes": (_sz, [C.POINTER(BranchformerLayerWeights), _i, _i, _i, _i, _i]),
    "smx_branchformer_encoder_fwd": (_i, [C.POINTER(BranchformerLayerWeights), _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp,
                                          _vp, _vp, _sz, _vp]),
    "smx_fbank_frames": (C.c_int32, [C.POINTER(FbankDesc), _i]),
    "smx_fbank_workspace_bytes": (_sz, [C.POINTER(FbankDesc), _i]),
    "smx_fbank_fwd": (_i, [C.POINTER(FbankDesc), _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "smx_input_norm_fwd": (_i, [_i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "smx_spec_drop_workspace_bytes": (_sz, []),
    "smx_spec_drop_fwd": (_i, [_i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _sz, _vp]),
    "smx_time_warp_fwd": (_i, [_i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "smx_conv_frontend_block_fwd": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "smx_input_proj_workspace_bytes": (_sz, [_i, _i, _i]),
    "smx_input_proj_fwd": (_i, [C.POINTER(Linear), _i, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "smx_padding_mask_from_wav_len": (_i, [_vp, _i, _i, _vp, _vp]),
    "smx_chunk_mask": (_i, [_i, _i, _i, _vp, _vp]),
    "smx_debug_tc_gemm": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "smx_debug_set_trace": (_i, [_vp]),
    "smx_debug_set_ffn_cluster": (_i, [_i]),
    "smx_debug_set_ffn_version": (_i, [_i]),
    "smx_debug_set_pdl": (_i, [_i]),
    "smx_debug_set_cell_version": (_i, [_i]),
    "smx_debug_set_f32_tc": (_i, [_i]),
}


ABI_STRUCTS = [Linear, CellWeights, FFNWeights, ConvModWeights, ConformerLayerWeights, ConvBranchWeights,
               BranchformerLayerWeights, CellGrads, FFNGrads, ConvModGrads, ConvBranchGrads]


def exported_symbols():
    """Names include/smx.h declares (tests check the built library exports every one)."""
    return sorted(_PROTOS)


def lib():
    """Load libsmx.so (built in-tree by summarymixing_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libsmx.so not found at {LIB_PATH}: build it with `python -m summarymixing_b200.build` "
                "(there is no CPU fallback)"
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise SmxError(code, lib().smx_last_error().decode())
