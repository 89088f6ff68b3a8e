"""Shared parity checks: per-QP relative l2 error against reference outputs.

Stated tolerances (fp64, SURVEY.md section 8c / BASELINE.md section 5):
z*, nu <= 1e-8 and every gradient <= 1e-6, per-QP relative l2; lambda and the slacks are compared entry-wise
with rtol 1e-6 and atol 1e-8 * max|ref| (entries on the inactive / active side are ~1e-20 .. 1e-12 and carry
no information: the backward pass clamps them at 1e-8, qp.py:148).

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


def check_against_golden(out, gold, full_mats, ztol=ZTOL, gtol=GTOL, what="", prob=None):
    """out: dict(zhat, lam, slacks, nus, grads=(dQ,dp,dG,dh,dA,db)).
    prob (optional, fully batched inputs): adds the natural-scale floor of `natural_grad_scales` to the gradient
    denominators (needed when EVERY QP of a case sits on a vertex, so that the batch maximum is clamp noise too)."""
    from oracle.cases import proj
    scales = None
    if prob is not None and np.asarray(prob["Q"]).ndim == 3 and prob.get("dl") is not None:
        scales = natural_grad_scales(prob, np.asarray(gold["zhat"]), np.asarray(gold["lam"]))
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
            e_rows = rel_rows(g, ref, floor=1e-4)
            if scales is not None and np.asarray(ref).ndim >= 2 and np.asarray(ref).shape[0] == len(scales[k]):
                r2 = np.asarray(ref, dtype=np.float64).reshape(len(scales[k]), -1)
                nr = np.linalg.norm(r2, axis=1)
                den = np.maximum(np.maximum(nr, 1e-4 * nr.max()), 1e-300)
                sc = scales[k] if k in gold else scales[k] * np.linalg.norm(proj(np.asarray(g).shape[-1])) / np.sqrt(r2.shape[1])
                e_rows = e_rows * den / np.maximum(den, 1e-5 * sc)
            errs[k] = e_rows.max()
            assert errs[k] <= gtol, (what, k, errs[k])
    return errs


# ---- the randomised sweep (oracle/cases.py: sweep_problem) ---------------------------------------------------
# The sweep contains problems on which the reference itself returns an inaccurate iterate (ill-conditioned Q with
# nearly every constraint active: its batch stops after three iterations in which no QP improved, batch.py:127-143,
# and prints INACC_ERR).  Equality with such an answer means nothing, so every QP is first classified with a
# solver-independent measure, the KKT residual of the REFERENCE's own (z, lambda, nu):
#   reference converged (kkt <= KKT_OK)  -> parity at the stated tolerances (scaled with cond(Q), see below);
#   reference did not converge           -> our KKT residual must not be worse than the reference's.
KKT_OK = 1e-6


def kkt_residual(prob, z, lam, nu):
    """max of relative stationarity, primal feasibility (ineq and eq), complementarity, dual sign; per QP."""
    Q, p, G, h, A, b = (np.asarray(prob[k], dtype=np.float64) for k in ("Q", "p", "G", "h", "A", "b"))
    out = []
    for i in range(z.shape[0]):
        Qz = Q[i] @ z[i]
        r = Qz + p[i] + G[i].T @ lam[i]
        if A.shape[1]:
            r = r + A[i].T @ nu[i]
        sc = max(1.0, np.linalg.norm(p[i]), np.linalg.norm(Qz))
        gz = G[i] @ z[i] - h[i]
        pri = max(0.0, gz.max()) / max(1.0, np.abs(h[i]).max())
        eqr = (np.abs(A[i] @ z[i] - b[i]).max() / max(1.0, np.abs(b[i]).max())) if A.shape[1] else 0.0
        comp = np.abs(lam[i] * gz).max() / sc
        out.append(max(np.linalg.norm(r) / sc, pri, eqr, comp, max(0.0, -lam[i].min())))
    return np.array(out)


def natural_grad_scales(prob, z, lam):
    """Per QP: the size each gradient would have if no inequality were active, N = |Q^-1 dl| times the factor its
    outer product carries (qp.py:157-176).  The reference's gradients carry absolute noise of ~1e-8 * N from its
    1e-8 clamps (qp.py:148): at a vertex solution the true dz*/dp is exactly 0 and the reference returns that
    noise.  Norms below 1e-5 * scale are therefore compared as if they were that large (absolute error 1e-11 N)."""
    Q, dl = np.asarray(prob["Q"], dtype=np.float64), np.asarray(prob["dl"], dtype=np.float64)
    B = z.shape[0]
    N = np.array([np.linalg.norm(np.linalg.solve(Q[i], dl[i])) for i in range(B)])
    zn = np.maximum(np.linalg.norm(z, axis=1), 1.0)
    ln = np.maximum(np.linalg.norm(lam, axis=1), 1.0)
    return dict(dQ=N * zn, dp=N, dG=N * np.maximum(zn, ln), dh=N, dA=N * np.maximum(zn, ln), db=N)


def sweep_errors(out, prob, gold):
    """Classify every QP of a sweep case and measure it. Returns dict of per-QP arrays / scalars (no asserts)."""
    B = gold["zhat"].shape[0]
    kr = kkt_residual(prob, gold["zhat"], gold["lam"], gold.get("nus"))
    ko = kkt_residual(prob, out["zhat"], out["lam"], out.get("nus"))
    conv = kr <= KKT_OK
    res = dict(ref_kkt=kr, our_kkt=ko, ref_converged=conv, z=rel_rows(out["zhat"], gold["zhat"]))
    # tolerance = the stated one, plus 3x what the reference's own output moves when its inputs are perturbed by
    # 1e-15 relative (golden `sens_*`, oracle/gen_golden.py); expressed in the same relative units
    res["ztol"], res["gtol"] = ZTOL, GTOL
    zn = np.maximum(np.linalg.norm(gold["zhat"], axis=1), 1e-300)
    # (per case: the largest sensitivity among its converged QPs - a single random perturbation can be lucky on one QP)
    zs = (gold["sens_zhat"] / zn)[conv].max() if ("sens_zhat" in gold and conv.any()) else 0.0
    res["z_allow"] = np.full(B, ZTOL + 3.0 * zs)
    if out.get("grads") is not None:
        sc = natural_grad_scales(prob, gold["zhat"], gold["lam"])
        for k, g in zip(("dQ", "dp", "dG", "dh", "dA", "db"), out["grads"]):
            if g is None or k not in gold:
                continue
            a = np.asarray(g).reshape(B, -1)
            r = np.asarray(gold[k]).reshape(B, -1)
            nr = np.linalg.norm(r, axis=1)
            den = np.maximum(np.maximum(nr, 1e-4 * nr[conv].max() if conv.any() else 0.0), 1e-5 * sc[k])
            den = np.maximum(den, 1e-300)
            res[k] = np.linalg.norm(a - r, axis=1) / den
            gs = (gold["sens_" + k] / den)[conv].max() if (("sens_" + k) in gold and conv.any()) else 0.0
            res[k + "_allow"] = GTOL + 3.0 * gs
    return res


def check_sweep(out, prob, gold, what=""):
    r = sweep_errors(out, prob, gold)
    conv = r["ref_converged"]
    if conv.any():
        assert (r["z"][conv] <= r["z_allow"][conv]).all(), (what, "z", r["z"], r["z_allow"])
        for k in ("dQ", "dp", "dG", "dh", "dA", "db"):
            if k in r:
                assert (r[k][conv] <= (r[k + "_allow"] * np.ones_like(r[k]))[conv]).all(), (what, k, r[k], r[k + "_allow"])
    if (~conv).any():     # the reference printed INACC_ERR here: be at least as close to a KKT point as it is
        assert (r["our_kkt"][~conv] <= np.maximum(r["ref_kkt"][~conv], KKT_OK)).all(), (what, r["our_kkt"], r["ref_kkt"])
    return r
