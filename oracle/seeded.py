"""Deterministic, seed-defined inputs and weights for the tile-aligned golden fixtures (tests/golden/tile/).

A D = 256 / 512 encoder has MBs of weights: instead of committing them, both the generator (oracle/gen_golden_tile.py,
which fills the UNMODIFIED reference modules) and the tests (which fill this repo's modules) derive every tensor from
numpy's PCG64 stream (stable across numpy versions by policy) keyed by (seed, crc32(parameter name)) -- independent of
parameter enumeration order.  Test infrastructure: imported only by tests/, oracle/ and bench.py's reference arm.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def seeded_normal(seed: int, name: str, shape) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64([int(seed), zlib.crc32(name.encode())]))
    return rng.standard_normal(tuple(shape), dtype=np.float64).astype(np.float32)


def seeded_param(seed: int, name: str, shape) -> torch.Tensor:
    """Weight-like values: matrices ~ N(0, 1/fan_in), LayerNorm gains ~ 1 + 0.1 N, every other vector ~ 0.1 N."""
    z = seeded_normal(seed, name, shape)
    leaf = name.rsplit(".", 1)[-1]
    if len(shape) >= 2:
        if leaf == "weights":                      # ParallelLinear (h, in/h, out/h)
            fan_in = shape[1]
        elif leaf == "biases":                     # ParallelLinear (h, out/h)
            return torch.from_numpy(0.1 * z)
        else:                                      # nn.Linear (out, in) / Conv1d (out, in/groups, k)
            fan_in = int(np.prod(shape[1:]))
        return torch.from_numpy(z / np.sqrt(float(fan_in)))
    if len(shape) == 0:
        raise ValueError(f"{name}: scalar parameters keep their constructor value")
    is_norm_gain = leaf == "weight" and ("norm" in name or name.split(".")[-2] in ("0",) or "layer_norm" in name)
    return torch.from_numpy(1.0 + 0.1 * z if is_norm_gain else 0.1 * z)


def fill_module(module: torch.nn.Module, seed: int) -> None:
    """Overwrite every parameter of `module` from the seeded stream (0-dim parameters, e.g. decay_constant, are kept)."""
    with torch.no_grad():
        for name, p in module.state_dict().items():
            if p.dim() == 0 or not p.dtype.is_floating_point:
                continue
            p.copy_(seeded_param(seed, name, p.shape).to(p.dtype))


def seeded_input(seed: int, B: int, T: int, D: int) -> torch.Tensor:
    """(B,T,D) standard-normal input, rounded to bf16-representable values so fp32 and bf16 arms see the same numbers."""
    x = torch.from_numpy(seeded_normal(seed, "input", (B, T, D)))
    return x.to(torch.bfloat16).float()
