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
