"""The fused OptNet layer (qpth_b200/layers.py) against (1) the notebook's block run with the REAL reference
(tests/golden/layer_*.npz, oracle/gen_golden_layer.py) and (2) the same formulas composed with torch autograd around
qpth_b200.QPFunction (isolates the construct / chain kernels: must agree to rounding)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b, floor=1e-4):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), floor * max(1.0, np.abs(b).max())))


@pytest.mark.parametrize("name", ["layer_small", "layer_cls"])
def test_fused_optnet_layer(name, golden_dir):
    from qpth_b200 import QPFunction
    from qpth_b200.layers import OptNetQP
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    t = {k: torch.tensor(z[k], dtype=torch.float64, device=DEV, requires_grad=True) for k in ("L", "G", "z0", "s0", "p")}
    dl = torch.tensor(z["dl"], dtype=torch.float64, device=DEV)
    eps = float(z["eps"])
    out = OptNetQP(eps=eps)(t["L"], t["G"], t["z0"], t["s0"], t["p"])
    out.backward(dl)
    got = {"z": out.detach().cpu().numpy(), **{"d" + k: t[k].grad.cpu().numpy() for k in ("L", "G", "z0", "s0", "p")}}
    assert _rel(got["z"], z["z"]) <= 1e-8
    for k in ("dL", "dG", "dz0", "ds0", "dp"):
        assert _rel(got[k], z[k]) <= 1e-6, (name, k, _rel(got[k], z[k]))
    # the same block composed with torch autograd around QPFunction
    u = {k: torch.tensor(z[k], dtype=torch.float64, device=DEV, requires_grad=True) for k in ("L", "G", "z0", "s0", "p")}
    n = u["L"].size(0)
    Lm = torch.tril(torch.ones(n, n, dtype=torch.float64, device=DEV)) * u["L"]
    Q = Lm.mm(Lm.t()) + eps * torch.eye(n, dtype=torch.float64, device=DEV)
    h = u["G"].mv(u["z0"]) + u["s0"]
    e = torch.Tensor().to(DEV).double()
    z2 = QPFunction(verbose=-1)(Q, u["p"], u["G"], h, e, e)
    z2.backward(dl)
    assert _rel(got["z"], z2.detach().cpu().numpy()) <= 1e-12
    for k in ("L", "G", "z0", "s0", "p"):
        assert _rel(got["d" + k], u[k].grad.cpu().numpy()) <= 1e-10, (name, k)
