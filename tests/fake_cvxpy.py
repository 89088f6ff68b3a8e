"""A test double for the few `cvxpy` names qpth's CVXPY front end uses (qpth/solvers/cvxpy.py:5-31): Variable, quad_form,
Minimize, Problem, affine expressions with `@ + ==  >=` and `dual_value`. The real package is not in this image, so
without it `qpth_b200.solution.cvxpy_forward` would never execute anywhere. `Problem.solve()` recognises exactly the
problem that function states - minimise 1/2 z'Qz + p'z s.t. Az = b, Gz + s = h, s >= 0 - and solves it with the numpy
oracle (one QP at a time). TEST INFRASTRUCTURE ONLY."""
import numpy as np


class _Affine:
    """sum_v M_v v + c over Variables."""
    __array_ufunc__ = None            # make `ndarray @ expr` / `ndarray + expr` defer to the reflected operators below

    def __init__(self, terms, const):
        self.terms, self.const = terms, const

    def __add__(self, other):
        if isinstance(other, _Quad):
            return other + self
        if isinstance(other, _Affine):
            t = dict(self.terms)
            for v, M in other.terms.items():
                t[v] = t[v] + M if v in t else M
            return _Affine(t, self.const + other.const)
        return _Affine(dict(self.terms), self.const + np.asarray(other, dtype=np.float64))
    __radd__ = __add__

    def __rmatmul__(self, M):
        M = np.asarray(M, dtype=np.float64)
        return _Affine({v: M @ T for v, T in self.terms.items()}, M @ self.const)

    def __eq__(self, rhs):            # noqa: PLW1641 (constraints, not equality)
        return _Constraint(self + (-np.asarray(rhs, dtype=np.float64)), "==")

    def __ge__(self, rhs):
        return _Constraint(self + (-np.asarray(rhs, dtype=np.float64)), ">=")
    __hash__ = object.__hash__


class Variable(_Affine):
    def __init__(self, n):
        self.n, self.value = n, None
        _Affine.__init__(self, {self: np.eye(n)}, np.zeros(n))
    __hash__ = object.__hash__
    __eq__ = _Affine.__eq__


class _Quad:
    __array_ufunc__ = None

    def __init__(self, var, Q, scale=1.0, lin=None):
        self.var, self.Q, self.scale, self.lin = var, np.asarray(Q, dtype=np.float64), scale, lin

    def __rmul__(self, a):
        return _Quad(self.var, self.Q, self.scale * float(a), self.lin)
    __mul__ = __rmul__

    def __add__(self, lin):
        assert self.lin is None and isinstance(lin, _Affine)
        return _Quad(self.var, self.Q, self.scale, lin)
    __radd__ = __add__


def quad_form(x, Q):
    assert isinstance(x, Variable)
    return _Quad(x, Q)


class _Constraint:
    def __init__(self, expr, kind):
        self.expr, self.kind, self.dual_value = expr, kind, None


class Minimize:
    def __init__(self, objective):
        self.objective = objective


class Problem:
    def __init__(self, objective, constraints):
        self.objective, self.constraints, self.status = objective.objective, constraints, None

    def solve(self):
        from oracle import pdipm_oracle as orc
        obj = self.objective
        z = obj.var
        nz = z.n
        Q = 2.0 * obj.scale * obj.Q                        # scale * z'Qz == 1/2 z'(2 scale Q) z
        p = np.asarray(obj.lin.terms[z]).reshape(nz)
        assert set(obj.lin.terms) == {z} and not np.any(obj.lin.const)
        eq = ineq = slack = None
        for c in self.constraints:
            if c.kind == ">=":                             # s >= 0
                (slack,) = c.expr.terms
                assert np.array_equal(c.expr.terms[slack], np.eye(slack.n)) and not np.any(c.expr.const)
            elif len(c.expr.terms) == 2:                   # G z + s - h == 0
                ineq = c
            else:                                          # A z - b == 0
                eq = c
        assert ineq is not None and slack is not None
        assert np.array_equal(ineq.expr.terms[slack], np.eye(slack.n))
        G, h = ineq.expr.terms[z], -ineq.expr.const
        if eq is not None:
            assert set(eq.expr.terms) == {z}
            A, b = eq.expr.terms[z], -eq.expr.const
        else:
            A, b = np.zeros((0,)), np.zeros((0,))
        r = orc.qp_solve(Q[None], p[None], G[None], h[None], A[None] if eq is not None else A,
                         b[None] if eq is not None else b, maxIter=40, per_qp=True)
        z.value, slack.value = r["zhat"][0], r["slacks"][0]
        ineq.dual_value = r["lam"][0]
        if eq is not None:
            eq.dual_value = r["nus"][0]
        self.status = "optimal"
        return 0.5 * z.value @ Q @ z.value + p @ z.value
