"""Shared parity checks: per-QP relative l2 error against reference outputs.

Stated tolerances (fp64, SURVEY.md section 8c / BASELINE.md section 5):
z*, lambda, s, nu <= 1e-8 and every gradient <= 1e-6, per-QP relative l2.

Gradients use a floored denominator: max(||ref_i||, 1e-4 * max_j ||ref_j||).  A QP whose
solution sits on a vertex (nz active constraints) has an exactly-zero dz*/dp; the
reference returns ~1e-8 noise there (an artefact of its 1e-8 clamp, qp.py:148), and a
pure relative error on noise is meaningless.
"""
import numpy as np

ZTOL = 1e-8
GTOL = 1e-6


def rel_rows(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.ndim <= 1:
        a = a.reshape(1, -1); b = b.reshape(1, -1)
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    nb = np.linalg.norm(b, axis=1)
    den = np.maximum(np.maximum(nb, floor * nb.max()), 1e-300)
    return np.linalg.norm(a - b, axis=1) / den


def check_against_golden(out, gold, full_mats, ztol=ZTOL, gtol=GTOL, what=""):
    """out: dict(zhat, lam, slacks, nus, grads=(dQ,dp,dG,dh,dA,db))."""
    from oracle.cases import proj
    errs = {}
    for k in ("zhat", "lam", "slacks", "nus"):
        if k in gold and out.get(k) is not None:
            errs[k] = rel_rows(out[k], gold[k]).max()
            # slacks/duals on the clamped side are ~1e-20 and only noise; compare absolutely too
            tol = ztol
            if k in ("lam", "slacks"):
                ok = np.allclose(out[k], gold[k], rtol=1e-6, atol=1e-8 * max(1.0, np.abs(gold[k]).max()))
                assert ok, (what, k, errs[k])
            else:
                assert errs[k] <= tol, (what, k, errs[k])
    if out.get("grads") is not None:
        for k, g in zip(("dQ", "dp", "dG", "dh", "dA", "db"), out["grads"]):
            if g is None:
                assert k not in gold and (k + "_proj") not in gold, (what, k)
                continue
            if k in gold:
                ref = gold[k]
            elif k + "_proj" in gold:
                ref = gold[k + "_proj"]
                g = np.asarray(g) @ proj(np.asarray(g).shape[-1])
            else:
                raise AssertionError((what, k, "missing in golden"))
            errs[k] = rel_rows(g, ref, floor=1e-4).max()
            assert errs[k] <= gtol, (what, k, errs[k])
    return errs
