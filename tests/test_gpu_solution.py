"""GPU parity of the backward-only entry (qpth/qp.py:97-120,142-182).

`QPSolutionFunction` is fed the REAL reference's primal/dual solution from tests/golden/ (so no PDIPM iterate of
ours is involved) and its gradients are compared with the reference's gradients for the same inputs, at the stated
gradient tolerance. A second test feeds it our own forward's solution: the gradients must then be the ones
`QPFunction` itself returns (same kernels, same inputs -> bitwise equal).
"""
import numpy as np
import pytest
import torch

from oracle.cases import load_case
from tests.parity import check_against_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tensors(prob):
    t = {}
    for k in ("Q", "p", "G", "h", "A", "b"):
        v = np.asarray(prob[k], dtype=np.float64)
        t[k] = torch.Tensor().to(DEV).double() if v.size == 0 else \
            torch.tensor(v, dtype=torch.float64, device=DEV, requires_grad=True)
    return t


@pytest.mark.parametrize("name", ["c1", "eq_small", "ineq_only_wide", "shared", "unbatched", "c4_small"])
def test_reference_solution_through_backward_only_entry(name, golden_dir):
    from qpth_b200 import QPSolutionFunction
    prob, gold, full = load_case(name, golden_dir)
    t = _tensors(prob)
    sol = [torch.tensor(np.asarray(gold[k]), dtype=torch.float64, device=DEV) for k in ("zhat", "lam", "slacks")]
    nus = torch.tensor(np.asarray(gold["nus"]), dtype=torch.float64, device=DEV) if "nus" in gold else \
        torch.Tensor().to(DEV).double()
    z = QPSolutionFunction()(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"], sol[0], sol[1], sol[2], nus)
    assert torch.equal(z.detach().view(-1), sol[0].view(-1))
    dl = torch.tensor(np.asarray(prob["dl"]).reshape(tuple(z.shape)), dtype=torch.float64, device=DEV)
    z.backward(dl)
    out = dict(grads=tuple(None if t[k].grad is None else t[k].grad.cpu().numpy()
                           for k in ("Q", "p", "G", "h", "A", "b")))
    check_against_golden(out, gold, full, what="solution:" + name)


def test_own_solution_gives_own_gradients():
    from qpth_b200 import QPFunction, QPSolutionFunction
    from qpth_b200.problems import random_qp_batch
    prob = random_qp_batch(16, 30, 20, 4, seed=11)
    t1, t2 = _tensors(prob), _tensors(prob)
    f = QPFunction(verbose=-1)
    z1 = f(t1["Q"], t1["p"], t1["G"], t1["h"], t1["A"], t1["b"])
    st = f.last_solve()
    dl = torch.tensor(prob["dl"], dtype=torch.float64, device=DEV)
    z1.backward(dl)
    z2 = QPSolutionFunction()(t2["Q"], t2["p"], t2["G"], t2["h"], t2["A"], t2["b"],
                              st.zhat, st.lam, st.slacks, st.nus)
    z2.backward(dl)
    for k in ("Q", "p", "G", "h", "A", "b"):
        assert torch.equal(t1[k].grad, t2[k].grad), k


def test_solution_shape_errors():
    from qpth_b200 import QPSolutionFunction
    from qpth_b200.problems import random_qp_batch
    prob = random_qp_batch(3, 6, 4, 0, seed=1)
    t = _tensors(prob)
    bad = torch.zeros(3, 5, dtype=torch.float64, device=DEV)
    ok = torch.zeros(3, 4, dtype=torch.float64, device=DEV)
    with pytest.raises(RuntimeError):
        QPSolutionFunction()(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"], bad, ok, ok, torch.Tensor())


def test_lower_triangle_strip_copies():
    """qpb200_copy_lower (util.copy_lower_): the lower triangle of a batch of symmetric matrices crosses in strips, in both
    directions; what lies above the strips in the destination is untouched."""
    import torch
    from qpth_b200.util import copy_lower_
    B, n = 5, 37
    h = torch.randn(B, n, n, dtype=torch.float64)
    h = (h + h.transpose(1, 2)).contiguous().pin_memory()
    for band in (8, 10, 37, 64):
        d = torch.full((B, n, n), -7.0, dtype=torch.float64, device="cuda:0")
        copy_lower_(d, h, band=band)
        torch.cuda.synchronize()
        dc = d.cpu()
        assert torch.equal(torch.tril(dc), torch.tril(h))
        r = torch.arange(n)
        beyond = (r[None, :] >= ((r[:, None] // band) + 1) * band)          # columns past the strip of each row
        assert bool((dc[:, beyond] == -7.0).all())
        back = torch.zeros(B, n, n, dtype=torch.float64).pin_memory()
        copy_lower_(back, d, band=band)
        torch.cuda.synchronize()
        assert torch.equal(torch.tril(back), torch.tril(h))


def test_lazy_checks_defer_the_flag_read(capsys):
    """qp.LAZY_CHECKS: the SPD / inaccuracy flags are examined later (flush_checks), not inside forward."""
    import torch
    from qpth_b200 import QPFunction, qp as qpmod
    from qpth_b200.problems import random_qp_batch
    pr = random_qp_batch(4, 12, 8, 0, seed=5)
    t = {k: torch.tensor(pr[k], dtype=torch.float64, device="cuda:0") for k in ("Q", "p", "G", "h")}
    e = torch.Tensor().to("cuda:0").double()
    bad = t["Q"].clone(); bad[1] = -bad[1]                     # not SPD
    old = qpmod.LAZY_CHECKS
    try:
        qpmod.LAZY_CHECKS = True
        QPFunction()(bad, t["p"], t["G"], t["h"], e, e)        # returns without raising
        with pytest.raises(RuntimeError, match="Q is not SPD"):
            qpmod.flush_checks()
        assert not qpmod._pending
        z = QPFunction()(t["Q"], t["p"], t["G"], t["h"], e, e)  # a good problem: nothing pending after the flush
        qpmod.flush_checks()
        assert bool(torch.isfinite(z).all())
        qpmod.LAZY_CHECKS = False
        with pytest.raises(RuntimeError, match="Q is not SPD"):
            QPFunction()(bad, t["p"], t["G"], t["h"], e, e)
    finally:
        qpmod.LAZY_CHECKS = old
        qpmod._pending.clear()
