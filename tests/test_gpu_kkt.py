"""The reference's stand-alone KKT solvers on the B200 kernels (qpth_b200/kkt.py) against outputs of the REAL
`factor_solve_kkt` / `solve_kkt_ir` (tests/golden/kkt_*.npz, oracle/gen_golden_kkt.py).

Tolerances: the exact solve (LU_FULL) <= 1e-8 per-QP relative. `solve_kkt_ir`: the reference's refinement step adds the
correction with the wrong sign (batch.py:258-261 passes -res to a routine that solves K v = -rhs), so its result carries
TWICE the O(eps) error of the regularised solve (residual ~1e-6) instead of O(eps^2); we compare to it at 1e-4 and
require our own residual to be smaller than the reference's."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["kkt_small", "kkt_c3", "kkt_ineq_only", "kkt_singular"]


def _load(name, golden_dir):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    t = {k: torch.tensor(z[k], dtype=torch.float64, device=DEV) for k in ("Q", "G", "A", "d", "rx", "rs", "rz", "ry")}
    neq = t["A"].shape[1]
    return z, t, (t["A"] if neq else None), (t["ry"] if neq else None)


def _rel(a, b):
    a, b = a.cpu().numpy(), np.asarray(b)
    return float((np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)).max())


@pytest.mark.parametrize("name", CASES)
def test_kkt_variants_vs_reference(name, golden_dir):
    from qpth_b200 import kkt
    z, t, A, ry = _load(name, golden_dir)
    D = torch.diag_embed(t["d"])
    if "full_dx" in z.files and name != "kkt_singular":
        out = kkt.factor_solve_kkt(t["Q"], D, t["G"], A, t["rx"], t["rs"], t["rz"], ry)
        for k, v in zip(("dx", "ds", "dz", "dy"), out):
            if v is not None:
                assert _rel(v, z["full_" + k]) <= 1e-8, (name, k)
    out = kkt.solve_kkt_ir(t["Q"], D, t["G"], A, t["rx"], t["rs"], t["rz"], ry, niter=1)
    assert all(bool(torch.isfinite(v).all()) for v in out if v is not None)
    res = kkt.kkt_resid_reg(t["Q"], t["d"], t["G"], A, kkt.IR_EPS, *out, t["rx"], t["rs"], t["rz"], ry)
    ours = max(float(v.abs().max()) for v in res if v is not None)
    assert ours <= float(z["ir_resid_max"]), (name, ours, float(z["ir_resid_max"]))
    if name != "kkt_singular":
        assert ours <= 1e-9, (name, ours)          # O(eps^2) after one correct refinement step
        for k, v in zip(("dx", "ds", "dz", "dy"), out):
            if v is not None and z["ir_" + k].size:
                assert _rel(v, z["ir_" + k]) <= 1e-4, (name, k)
    else:
        # two more steps: the refinement converges on the regularised-singular system as well
        out3 = kkt.solve_kkt_ir(t["Q"], D, t["G"], A, t["rx"], t["rs"], t["rz"], ry, niter=3)
        res3 = kkt.kkt_resid_reg(t["Q"], t["d"], t["G"], A, kkt.IR_EPS, *out3, t["rx"], t["rs"], t["rz"], ry)
        assert max(float(v.abs().max()) for v in res3 if v is not None) <= ours


def test_plain_cholesky_path_fails_where_ir_survives(golden_dir):
    """PSD-singular Q: the un-regularised pre_factor_kkt reports a failed pivot (the reason the variant exists)."""
    from qpth_b200 import kkt
    z, t, A, ry = _load("kkt_singular", golden_dir)
    f = kkt._Factored(t["Q"], t["G"], A, 0.0)
    torch.cuda.synchronize()
    assert int(f.spd.sum()) > 0 or not bool(torch.isfinite(f.K).all())
    g = kkt._Factored(t["Q"], t["G"], A, kkt.IR_EPS)
    torch.cuda.synchronize()
    assert int(g.spd.sum()) == 0 and bool(torch.isfinite(g.K).all())
