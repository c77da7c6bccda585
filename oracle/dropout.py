"""TEST INFRASTRUCTURE (CPU oracle; never imported by the product package): the counter-based dropout masks of libsmx's
training path (include/smx.h, smx_dropout), restated in numpy integer arithmetic, and a hook that applies them inside the
oracle's forward functions (oracle/smx_oracle.py, `drop=`).

The reference applies torch.nn.Dropout (summary_mixing.py:114,252,297; Conformer.py:163,470-484), whose random stream
libsmx does not reproduce; what is pinned here is (a) the mask generator, bit-exactly (Philox4x32-10: Salmon et al., SC'11,
constants 0xD2511F53 / 0xCD9E8D57, Weyl keys 0x9E3779B9 / 0xBB67AE85 — checked against the published known-answer vectors in
tests/test_dropout_cpu.py) and (b) the reference algorithm evaluated with those masks in place of torch's.

Element e of site s under (p, seed): counter = (lo32(e // 4), hi32(e // 4), s, 0), key = (lo32(seed), hi32(seed)),
word = e % 4; the element is dropped when word < floor(p * 2^32) (capped at 2^32 - 1), kept values are scaled by 1 / (1 - p)."""
from __future__ import annotations

import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
LO = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Ten rounds of Philox-4x32 on arrays of 32-bit words held in uint64."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2  # 32 x 32 -> 64 bit products
        hi0, lo0, hi1, lo1 = p0 >> S32, p0 & LO, p1 >> S32, p1 & LO
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & LO, (k1 + W1) & LO
    return c0, c1, c2, c3


def threshold(p: float) -> int:
    t = float(np.float32(p)) * 4294967296.0  # p travels as a C float
    return 4294967295 if t >= 4294967295.0 else int(t)


def keep_mask(p: float, seed: int, site: int, n: int) -> np.ndarray:
    """uint8 (n,): 1 where element e of `site` survives."""
    quads = np.arange((n + 3) // 4, dtype=np.uint64)
    z = np.zeros_like(quads)
    w = philox4x32_10(quads & LO, quads >> S32, z + np.uint64(site), z, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=1).reshape(-1)[:n]
    return (words >= np.uint64(threshold(p))).astype(np.uint8)


class Hook:
    """drop(key, tensor) for the oracle's forward functions: `sites` maps a key (the oracle's name of a dropout call:
    '<ffn prefix>inner' / '<ffn prefix>outer', '<cell prefix>cat', '<conv prefix>out') to (seed, site)."""

    def __init__(self, p: float, sites: dict):
        self.p, self.sites = float(np.float32(p)), sites
        self.used = []

    def __call__(self, key: str, t: torch.Tensor) -> torch.Tensor:
        seed, site = self.sites[key]
        self.used.append(key)
        keep = torch.from_numpy(keep_mask(self.p, seed, site, t.numel())).reshape(t.shape)
        return t * keep.to(t.dtype) * (1.0 / (1.0 - self.p))
