"""CPU-side checks of the backward-only entry: API surface and loud failures (no GPU, no cvxpy in this image)."""
import pytest
import torch


def test_exports():
    import qpth_b200
    assert callable(qpth_b200.QPSolutionFunction)
    assert qpth_b200.QPSolvers.CVXPY.value == 2        # qpth/qp.py:13-15


def test_cvxpy_branch_fails_loudly_without_cvxpy():
    pytest.importorskip("torch")
    try:
        import cvxpy  # noqa: F401
        pytest.skip("cvxpy is installed here")
    except ImportError:
        pass
    from qpth_b200 import QPFunction, QPSolvers
    f = QPFunction(solver=QPSolvers.CVXPY)
    Q = torch.eye(3, dtype=torch.float64).unsqueeze(0)
    p = torch.zeros(1, 3, dtype=torch.float64)
    G = torch.ones(1, 2, 3, dtype=torch.float64)
    h = torch.ones(1, 2, dtype=torch.float64)
    e = torch.Tensor()
    with pytest.raises(ImportError, match="cvxpy"):
        f(Q, p, G, h, e, e)


def test_backward_only_entry_needs_a_device():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from qpth_b200 import QPSolutionFunction, _lib
    Q = torch.eye(3, dtype=torch.float64).unsqueeze(0)
    p = torch.zeros(1, 3, dtype=torch.float64)
    G = torch.ones(1, 2, 3, dtype=torch.float64)
    h = torch.ones(1, 2, dtype=torch.float64)
    e = torch.Tensor()
    z = torch.zeros(1, 3, dtype=torch.float64)
    l = torch.zeros(1, 2, dtype=torch.float64)
    with pytest.raises(_lib.QpthB200Error):
        QPSolutionFunction()(Q, p, G, h, e, e, z, l, l, e)


@pytest.mark.parametrize("neq", [0, 2])
def test_cvxpy_front_end_runs_against_a_test_double(neq, monkeypatch):
    """`cvxpy_forward` (the reference's per-sample CVXPY solve, qp.py:97-120 / solvers/cvxpy.py:5-31) stated against
    tests/fake_cvxpy.py: the problem it builds must be THE QP (checked by the double and by the KKT conditions of what
    comes back), and the four results must come back in the reference's order (zhats, nus, lams, slacks)."""
    import sys
    import numpy as np
    try:
        import cvxpy  # noqa: F401
        pytest.skip("the real cvxpy is installed; this test drives the test double")
    except ImportError:
        pass
    from tests import fake_cvxpy
    from qpth_b200 import cvxpy_forward
    monkeypatch.setitem(sys.modules, "cvxpy", fake_cvxpy)
    rs = np.random.RandomState(5 + neq)
    B, nz, nineq = 3, 6, 4
    L = rs.randn(B, nz, nz)
    Q = L @ L.transpose(0, 2, 1) + 1e-2 * np.eye(nz)
    p = rs.randn(B, nz)
    G = rs.randn(B, nineq, nz)
    z0 = rs.randn(B, nz)
    h = np.einsum("bij,bj->bi", G, z0) + rs.rand(B, nineq)
    A = rs.randn(B, neq, nz)
    b = np.einsum("bij,bj->bi", A, z0)
    t = lambda x: torch.from_numpy(x)
    At, bt = (t(A), t(b)) if neq else (torch.Tensor(), torch.Tensor())
    zhats, nus, lams, slacks = cvxpy_forward(t(Q), t(p), t(G), t(h), At, bt)
    assert zhats.shape == (B, nz) and lams.shape == (B, nineq) and slacks.shape == (B, nineq)
    assert nus.shape == ((B, neq) if neq else (0,))
    z, lam, s = zhats.numpy(), lams.numpy(), slacks.numpy()
    stat = np.einsum("bij,bj->bi", Q, z) + p + np.einsum("bji,bj->bi", G, lam)
    if neq:
        stat = stat + np.einsum("bji,bj->bi", A, nus.numpy())
        assert np.abs(np.einsum("bij,bj->bi", A, z) - b).max() < 1e-9
    assert np.abs(stat).max() < 1e-8                                         # stationarity
    assert np.abs(np.einsum("bij,bj->bi", G, z) + s - h).max() < 1e-9        # G z + s = h
    assert (s > 0).all() and (lam > 0).all() and np.abs(s * lam).max() < 1e-8  # complementary slackness
