"""Host logic of the equality-only extension (qpth_b200/eqonly.py) on the CPU: the two stand-alone KKT solves are
replaced by a dense numpy solve of the SAME system the kernels solve (batch.py:313-346 sign convention), so signs,
argument order, gradient formulas (qp.py:148-177) and the `.mean(0)` rule are checked against torch autograd through
the closed-form KKT solution. The kernels behind the real `_Factored` are covered by tests/test_gpu_kkt.py and, for
this entry, by the equality-only jobs of tests/test_gpu_zz_fallback_families.py."""
import numpy as np
import pytest
import torch


class DenseFactor:
    """Stand-in for qpth_b200.kkt._Factored: [Q 0 G' A'; 0 D I 0; G I 0 0; A 0 0 0] [dx ds dz dy] = -[rx rs rz ry]."""

    def __init__(self, Q, G, A, reg):
        assert reg == 0.0
        self.Q, self.G, self.A = (t.detach().cpu().numpy() for t in (Q, G, A))
        self.spd = torch.tensor([int(np.linalg.eigvalsh(q).min() <= 0) for q in self.Q])

    def solve(self, d, rx, rs, rz, ry):
        d, rx, rs, rz, ry = (t.detach().cpu().numpy() for t in (d, rx, rs, rz, ry))
        B, m, n = self.G.shape
        e = self.A.shape[1]
        out = []
        for i in range(B):
            K = np.zeros((n + 2 * m + e,) * 2)
            K[:n, :n] = self.Q[i]; K[:n, n + m:n + 2 * m] = self.G[i].T; K[:n, n + 2 * m:] = self.A[i].T
            K[n:n + m, n:n + m] = np.diag(d[i]); K[n:n + m, n + m:n + 2 * m] = np.eye(m)
            K[n + m:n + 2 * m, :n] = self.G[i]; K[n + m:n + 2 * m, n:n + m] = np.eye(m)
            K[n + 2 * m:, :n] = self.A[i]
            out.append(np.linalg.solve(K, -np.concatenate([rx[i], rs[i], rz[i], ry[i]])))
        s = torch.tensor(np.stack(out))
        return s[:, :n], s[:, n:n + m], s[:, n + m:n + 2 * m], s[:, n + 2 * m:]


def _closed_form(Q, p, A, b):
    B, e, n = A.shape
    K = torch.cat([torch.cat([Q, A.transpose(1, 2)], 2), torch.cat([A, torch.zeros(B, e, e, dtype=Q.dtype)], 2)], 1)
    return torch.linalg.solve(K, torch.cat([-p, b], 1).unsqueeze(-1)).squeeze(-1)[:, :n]


@pytest.fixture
def on_cpu(monkeypatch):
    from qpth_b200 import eqonly
    monkeypatch.setattr(eqonly, "_factor", DenseFactor)
    monkeypatch.setattr(eqonly, "_target_device", lambda Q_: torch.device("cpu"))
    return eqonly


@pytest.mark.parametrize("shared", [False, True])
def test_equality_only_matches_the_closed_form(on_cpu, shared):
    from qpth_b200 import QPFunction
    rs = np.random.RandomState(3)
    B, nz, neq = 4, 9, 4
    L = rs.randn(nz, nz) if shared else rs.randn(B, nz, nz)
    Q = torch.tensor(L @ np.swapaxes(L, -1, -2) + 0.1 * np.eye(nz), requires_grad=True)
    p = torch.tensor(rs.randn(B, nz), requires_grad=True)
    A = torch.tensor(rs.randn(neq, nz) if shared else rs.randn(B, neq, nz), requires_grad=True)
    b = torch.tensor(rs.randn(B, neq), requires_grad=True)
    dl = torch.tensor(rs.randn(B, nz))
    e = torch.Tensor()
    z = QPFunction(verbose=-1)(Q, p, e, e, A, b)
    z.backward(dl)
    got = [t.grad.clone() for t in (Q, p, A, b)]
    for t in (Q, p, A, b):
        t.grad = None
    Qb = Q.expand(B, nz, nz) if shared else Q
    Ab = A.expand(B, neq, nz) if shared else A
    zc = _closed_form(Qb, p, Ab, b)
    assert torch.allclose(z, zc, rtol=1e-10, atol=1e-12)
    zc.backward(dl)
    # the reference's conventions: dQ symmetrised (qp.py:175-177); un-batched inputs get the batch MEAN (qp.py:159-177),
    # where autograd through expand() gives the batch SUM
    dQ = 0.5 * (Q.grad + Q.grad.transpose(-1, -2)) / (B if shared else 1)
    dA = A.grad / (B if shared else 1)
    for g, r in zip(got, (dQ, p.grad, dA, b.grad)):
        assert g.shape == r.shape and torch.allclose(g, r, rtol=1e-9, atol=1e-11)


def test_equality_only_checks_spd_and_shapes(on_cpu):
    from qpth_b200 import QPFunction
    e = torch.Tensor()
    Q = -torch.eye(3, dtype=torch.float64).unsqueeze(0)
    A = torch.ones(1, 1, 3, dtype=torch.float64); b = torch.ones(1, 1, dtype=torch.float64)
    p = torch.zeros(1, 3, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="Q is not SPD."):
        QPFunction()(Q, p, e, e, A, b)
    with pytest.raises(RuntimeError, match="inconsistent shapes"):
        QPFunction()(-Q, p, e, e, torch.ones(1, 1, 4, dtype=torch.float64), b)


def test_without_any_constraint_the_reference_assert_stays():
    from qpth_b200 import QPFunction
    e = torch.Tensor()
    with pytest.raises((AssertionError, RuntimeError)):
        QPFunction()(torch.eye(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64), e, e, e, e)
