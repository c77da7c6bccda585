"""CPU: the oracle's restatement of libsmx's counter-based dropout masks (oracle/dropout.py) against the published Philox4x32-10
known-answer vectors (Random123 kat_vectors), the mask statistics, and the oracle's `drop=` hook."""
import numpy as np
import torch

from oracle import dropout as OD
from oracle import smx_oracle as O


def test_philox4x32_10_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = OD.philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(v[0]) for v in got) == want


def test_keep_mask_is_a_function_of_seed_site_and_index():
    a = OD.keep_mask(0.3, 1234567890123, 1, 100003)
    assert a.dtype == np.uint8 and a.shape == (100003,)
    assert abs(a.mean() - 0.7) < 0.01
    np.testing.assert_array_equal(a[:1001], OD.keep_mask(0.3, 1234567890123, 1, 1001))  # a prefix does not depend on n
    assert (a != OD.keep_mask(0.3, 1234567890123, 0, 100003)).mean() > 0.3               # another site: another mask
    assert (a != OD.keep_mask(0.3, 1234567890124, 1, 100003)).mean() > 0.3               # another seed
    assert OD.keep_mask(0.0, 5, 0, 64).all()
    assert OD.threshold(0.5) == 1 << 31


def test_oracle_hook_scales_kept_values_and_names_every_site():
    torch.manual_seed(0)
    import summarymixing_b200 as S

    m = S.ConformerEncoderLayer(32, 64, 2, 7, attention_type="SummaryMixing", local_proj_hid_dim=[32], local_proj_out_dim=32,
                                summary_hid_dim=[32], dropout=0.1)
    sd = dict(m.state_dict())
    x = torch.randn(2, 20, 32)
    sites = {"ffn_module1.inner": (1, 0), "ffn_module1.outer": (1, 1), "mha_layer.cat": (2, 0), "convolution_module.out": (3, 0),
             "ffn_module2.inner": (4, 0), "ffn_module2.outer": (4, 1)}
    hook = OD.Hook(0.1, sites)
    y = O.conformer_layer(x, sd, "", act="swish", drop=hook)
    assert sorted(hook.used) == sorted(sites)
    y0 = O.conformer_layer(x, sd, "", act="swish")
    assert float((y - y0).abs().max()) > 1e-3
    t = torch.ones(4, 250)
    d = OD.Hook(0.2, {"k": (9, 0)})("k", t)
    assert set(torch.unique(d).tolist()) == {0.0, 1.25}
