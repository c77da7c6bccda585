"""CPU checks of the drop-in boundary: module surface, state_dict compatibility with the reference
(through the golden fixtures' state_dicts), the C-ABI library's exports and struct layout, and the
loud failure when no CUDA device is given.  No compute call is made here."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

import summarymixing_b200 as S
from summarymixing_b200 import _lib as L
from tests import _golden as G
from tests._build import module_from_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from summarymixing_b200 import build

    path = build.build()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "smx.h")).read()
    declared = sorted(set(re.findall(r"SMX_API\s+[\w\s\*]+?\b(smx_\w+)\s*\(", header)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/smx.h but not exported"
    assert declared == L.exported_symbols(), "ctypes binding and header disagree"
    assert L.lib().smx_version() == 100


def test_ctypes_structs_match_compiled_layout():
    lib = L.lib()
    for i, st in enumerate(L.ABI_STRUCTS):
        assert lib.smx_struct_size(i) == ctypes.sizeof(st), st.__name__


@pytest.mark.parametrize("name", G.names())
def test_reference_state_dict_loads_strict(name):
    fx = G.Fixture(name)
    m = module_from_fixture(fx)  # load_state_dict(strict=True) inside
    sd = m.state_dict()
    assert set(sd.keys()) == set(fx.sd.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(fx.sd[k].shape), k


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError, match="The SummaryMixing mode should either be"):
        S.SummaryMixing(64, 1, mode="nope")
    with pytest.raises(ValueError, match="dividible by n_split"):
        S.ParallelLinear(10, input_size=64, n_split=4)
    with pytest.raises(ValueError, match="Expected one of input_shape or input_size"):
        S.ParallelLinear(8)
    with pytest.raises(ValueError, match="must match dnn_blocks"):
        S.VanillaNN([None, None, 8], dnn_blocks=2, dnn_neurons=[4])


def test_reference_unit_test_shapes_construct():
    # tests/unittests/test_summary_mixing.py:17-54 of the reference: these four constructions must work
    for mode in ("SummaryMixing", "SummaryMixing-lite"):
        for nhead in (1, 4):
            S.SummaryMixing(enc_dim=64, nhead=nhead, local_proj_hid_dim=[32], local_proj_out_dim=32,
                            summary_out_dim=64, mode=mode)


def test_cpu_input_fails_loudly():
    sm = S.SummaryMixing(64, 4, [64], 64, [64], 64).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        sm(torch.randn(2, 5, 64))
    enc = S.ConformerEncoder(1, 64, 128, 4, attention_type="SummaryMixing", local_proj_hid_dim=[64],
                             local_proj_out_dim=64, summary_hid_dim=[64]).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.randn(2, 40, 64))


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        S.ConformerEncoderLayer(64, 128, 4, attention_type="RelPosMHAXL")
    with pytest.raises(NotImplementedError):
        S.SummaryMixing(64, 1, activation=nn.Softplus)


def test_product_package_never_imports_the_oracle():
    import glob

    for f in glob.glob(os.path.join(ROOT, "summarymixing_b200", "**", "*.py"), recursive=True):
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+\.*oracle|smx_oracle|sbshim", src, flags=re.M), f


def test_invalidate_weights_drops_every_cache():
    """summarymixing_b200.invalidate_weights: every module holding a weight cache forgets its key (the next call re-reads the
    parameters and re-packs the tensor-core images): the escape hatch for writes that bypass the version counters (p.data...)."""
    import summarymixing_b200 as S

    enc = S.ConformerEncoder(2, 64, 128, 4, 31, attention_type="SummaryMixing", local_proj_hid_dim=[64], local_proj_out_dim=64,
                             summary_hid_dim=[64])
    holders = [m for m in enc.modules() if hasattr(m, "_wv")]
    assert len(holders) >= 5
    for m in holders:
        m._wv._key = ("stale",)
    assert S.invalidate_weights(enc) == len(holders)
    assert all(m._wv._key is None for m in holders)
