"""The oracle's backward (torch.autograd through oracle/smx_oracle.py) pinned against gradients the UNMODIFIED
reference produced (tests/golden/bwd/*.npz, written by oracle/gen_golden_bwd.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from tests import _golden as G

BWD_DIR = os.path.join(G.GOLDEN_DIR, "bwd")


def bwd_names():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(BWD_DIR, "*.npz")))


def load_bwd(name):
    z = np.load(os.path.join(BWD_DIR, name + ".npz"))
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad.")}
    return torch.from_numpy(z["dy"]), torch.from_numpy(z["dx"]), grads


def oracle_grads(fx, dy, dtype=torch.float32):
    """dx and parameter gradients of the oracle on the fixture's input (autograd on the CPU)."""
    import copy

    from tests.test_oracle_golden import run_oracle

    fx = copy.copy(fx)
    fx.sd = {k: v.to(dtype).clone().requires_grad_(v.is_floating_point()) for k, v in fx.sd.items()}
    fx.x = fx.x.to(dtype).clone().requires_grad_(True)
    y = run_oracle(fx, dtype)
    y.backward(dy.to(dtype))
    return fx.x.grad, {k: v.grad for k, v in fx.sd.items() if v.grad is not None}


def test_fixtures_present():
    assert len(bwd_names()) >= 25


@pytest.mark.parametrize("name", bwd_names())
def test_oracle_autograd_matches_reference_gradients(name):
    fx = G.Fixture(name)
    dy, dx_ref, g_ref = load_bwd(name)
    dx, g = oracle_grads(fx, dy)
    assert float((dx - dx_ref).abs().max()) <= 2e-5 * max(1.0, float(dx_ref.abs().max()))
    assert set(g_ref) <= set(g)
    for k, v in g_ref.items():
        assert float((g[k] - v).abs().max()) <= 2e-5 * max(1.0, float(v.abs().max())), k
