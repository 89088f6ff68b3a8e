"""Run the REAL reference (locuslab/qpth at /root/reference) on CPU (TEST INFRASTRUCTURE ONLY).

Only usable in the build container where /root/reference exists; the GPU box
never imports this.  `import qpth` needs cvxpy (absent) only for its optional
CVXPY solver branch (`qpth/solvers/__init__.py:3`), so an empty stub module is
injected first — the PDIPM path runs unmodified.
"""
import os
import sys
import types
import warnings

REF = os.environ.get("QPTH_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "qpth"))


def load():
    if "cvxpy" not in sys.modules:
        sys.modules["cvxpy"] = types.ModuleType("cvxpy")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import qpth.qp as ref_qp          # noqa: E402
        import qpth.solvers.pdipm.batch as ref_batch   # noqa: E402
    return ref_qp, ref_batch


def run_reference(prob, requires=("Q", "p", "G", "h", "A", "b"), threads=None, **opts):
    """prob: dict of numpy arrays (Q,p,G,h,A,b,dl). Returns dict of numpy outputs.

    Empty A/b (neq == 0) are passed as `torch.Tensor()` exactly as the
    reference's callers do (`prof-linear.py:86`).
    """
    import numpy as np
    import torch
    if threads:
        torch.set_num_threads(threads)
    ref_qp, _ = load()
    t = {}
    for k in ("Q", "p", "G", "h", "A", "b"):
        v = np.asarray(prob[k], dtype=np.float64)
        t[k] = torch.from_numpy(v.copy()) if v.size > 0 else torch.DoubleTensor()
        if k in requires and v.size > 0:
            t[k].requires_grad_(True)
    captured = {}
    f = ref_qp.QPFunction(verbose=-1, **opts)
    z = f(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"])
    # ctx is not reachable from outside autograd; re-derive duals through a hook-free path:
    out = dict(zhat=z.detach().numpy().copy())
    if prob.get("dl") is not None:
        dl = torch.from_numpy(np.asarray(prob["dl"], dtype=np.float64).reshape(tuple(z.shape)).copy())
        z.backward(dl)
        for k in ("Q", "p", "G", "h", "A", "b"):
            g = t[k].grad
            out["d" + k] = None if g is None else g.numpy().copy()
    return out


def run_reference_duals(prob, **opts):
    """Forward only via the reference's own pre_factor_kkt + forward: returns x, y, z, s."""
    import numpy as np
    import torch
    ref_qp, ref_batch = load()
    from qpth.util import expandParam, extract_nBatch
    t = []
    for k in ("Q", "p", "G", "h", "A", "b"):
        v = np.asarray(prob[k], dtype=np.float64)
        t.append(torch.from_numpy(v.copy()) if v.size > 0 else torch.DoubleTensor())
    nB = extract_nBatch(*t)
    Q, _ = expandParam(t[0], nB, 3); p, _ = expandParam(t[1], nB, 2)
    G, _ = expandParam(t[2], nB, 3); h, _ = expandParam(t[3], nB, 2)
    A, _ = expandParam(t[4], nB, 3); b, _ = expandParam(t[5], nB, 2)
    Q_LU, S_LU, R = ref_batch.pre_factor_kkt(Q, G, A)
    x, y, z, s = ref_batch.forward(Q, p, G, h, A, b, Q_LU, S_LU, R,
                                   opts.get("eps", 1e-12), -1,
                                   opts.get("notImprovedLim", 3), opts.get("maxIter", 20))
    return dict(zhat=x.numpy().copy(), nus=None if y is None else y.numpy().copy(),
                lam=z.numpy().copy(), slacks=s.numpy().copy())
