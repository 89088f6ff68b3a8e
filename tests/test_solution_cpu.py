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
