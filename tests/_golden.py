"""Loader for tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_inputs = None


def names(prefix=""):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz"))):
        n = os.path.basename(f)[:-4]
        if n not in ("_inputs", "masks"):
            out.append(n)
    return out


class Fixture:
    def __init__(self, name):
        global _inputs
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        if _inputs is None:
            _inputs = np.load(os.path.join(GOLDEN_DIR, "_inputs.npz"))
        self.name = name
        self.cfg = json.loads(bytes(z["cfg"]).decode())
        self.y = torch.from_numpy(z["y"])
        self.x = torch.from_numpy(_inputs[bytes(z["in.x_ref"]).decode()])
        self.mask = torch.from_numpy(z["in.mask"]) if "in.mask" in z.files else None
        self.sum_mask = torch.from_numpy(z["in.sum_mask"]) if "in.sum_mask" in z.files else None
        self.sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}


# ---- tile-aligned fixtures (tests/golden/tile/, oracle/gen_golden_tile.py): seed-defined weights and inputs -------------
TILE_DIR = os.path.join(GOLDEN_DIR, "tile")


def tile_names(prefix=""):
    return [os.path.basename(f)[:-4] for f in sorted(glob.glob(os.path.join(TILE_DIR, prefix + "*.npz")))]


class TileFixture:
    """y32: the reference's fp32 output; y16: the reference module run in bfloat16 (its own bf16-vs-fp32 error is
    |y16 - y32|); x and the weights are regenerated from the seeds in cfg (oracle/seeded.py)."""

    def __init__(self, name):
        from oracle.seeded import seeded_input

        z = np.load(os.path.join(TILE_DIR, name + ".npz"))
        self.name = name
        self.cfg = json.loads(bytes(z["cfg"]).decode())
        self.y32 = torch.from_numpy(z["y32"])
        self.y16 = torch.from_numpy(z["y16"]).view(torch.bfloat16).float()
        self.mask = torch.from_numpy(z["mask"])
        c = self.cfg
        D = c.get("enc_dim") or c.get("d_model") or c.get("input_size")
        self.x = seeded_input(c["seed_x"], c["B"], c["T"], D)

    def ref_bf16_error(self):
        d = self.y16 - self.y32
        return float(d.abs().max()), float(d.norm() / self.y32.norm())
