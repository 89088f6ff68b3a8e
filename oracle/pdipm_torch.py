"""Batched CPU port of the reference's PDIPM forward + backward on torch CPU tensors
(TEST / BASELINE INFRASTRUCTURE ONLY — never imported by the product).

Why a second restatement: `pdipm_oracle.py` (numpy/scipy, per-QP loops) is the readable
checker but is slow; this one is the CPU BASELINE `bench.py` times (`cpu_baseline`,
`--impl reference`): the same algorithm as qpth/solvers/pdipm/batch.py:47-207,349-470
and qpth/qp.py:128-182 with batch-global semantics, expressed with the batched LAPACK
calls the reference itself uses (torch.linalg.lu_factor / lu_solve / bmm, all host
threads), minus the reference's `lu_unpack` re-pivot glue (batch.py:450-467) — the block
LU is applied as block elimination instead, so this port is, if anything, faster than
the reference on the same cores.  Checked against the golden vectors in tests/test_oracle.py.
"""
import torch


def _lu(x):
    return torch.linalg.lu_factor(x)           # pivoted on CPU, raises on an exactly singular matrix


def _lus(f, rhs):
    return torch.linalg.lu_solve(f[0], f[1], rhs)


class Factors:
    """pre_factor_kkt (batch.py:375-429)."""

    def __init__(self, Q, G, A):
        self.e = A.size(1) if A.nelement() > 0 else 0
        self.Q_LU = _lu(Q)
        GT = G.transpose(1, 2)
        self.R = G.bmm(_lus(self.Q_LU, GT))
        if self.e > 0:
            invQ_AT = _lus(self.Q_LU, A.transpose(1, 2))
            self.A_LU = _lu(A.bmm(invQ_AT))
            self.GiA = G.bmm(invQ_AT)                                   # G Q^-1 A^T
            self.R = self.R - self.GiA.bmm(_lus(self.A_LU, self.GiA.transpose(1, 2)))
        self.T_LU = None

    def factor_kkt(self, d):
        """batch.py:435-470."""
        T = self.R.clone()
        T.diagonal(dim1=1, dim2=2).add_(1.0 / d)
        self.T_LU = _lu(T)

    def solve_kkt(self, G, A, d, rx, rs, rz, ry):
        """batch.py:349-372."""
        invQ_rx = _lus(self.Q_LU, rx.unsqueeze(2))
        hz = G.bmm(invQ_rx).squeeze(2) + rs / d - rz
        if self.e > 0:
            hy = A.bmm(invQ_rx).squeeze(2) - ry
            u = _lus(self.A_LU, (-hy).unsqueeze(2))
            wz = _lus(self.T_LU, (-hz).unsqueeze(2) - self.GiA.bmm(u))
            wy = u - _lus(self.A_LU, self.GiA.transpose(1, 2).bmm(wz))
            g1 = -rx - G.transpose(1, 2).bmm(wz).squeeze(2) - A.transpose(1, 2).bmm(wy).squeeze(2)
            dy = wy.squeeze(2)
        else:
            wz = _lus(self.T_LU, (-hz).unsqueeze(2))
            g1 = -rx - G.transpose(1, 2).bmm(wz).squeeze(2)
            dy = None
        wz = wz.squeeze(2)
        dx = _lus(self.Q_LU, g1.unsqueeze(2)).squeeze(2)
        ds = (-rs - wz) / d
        return dx, ds, wz, dy


def get_step(v, dv):
    """batch.py:210-213 (batch-global fill value)."""
    a = -v / dv
    amax = a.max()
    fill = amax if bool(amax > 1.0) else torch.tensor(1.0, dtype=v.dtype)
    a = torch.where(dv > 0, fill, a)
    return a.min(1)[0]


def forward(Q, p, G, h, A, b, F, eps=1e-12, notImprovedLim=3, maxIter=20):
    """batch.py:47-207."""
    B, m, n = G.shape
    e = F.e
    d = torch.ones(B, m, dtype=Q.dtype)
    F.factor_kkt(d)
    zeros_m = torch.zeros(B, m, dtype=Q.dtype)
    x, s, z, y = F.solve_kkt(G, A, d, p, zeros_m, -h, -b if e > 0 else None)
    M = s.min(1)[0]
    s = torch.where((M < 0).unsqueeze(1), s - (M - 1).unsqueeze(1), s)
    M = z.min(1)[0]
    z = torch.where((M < 0).unsqueeze(1), z - (M - 1).unsqueeze(1), z)
    best = None
    nNot = 0
    for it in range(maxIter):
        rx = G.transpose(1, 2).bmm(z.unsqueeze(2)).squeeze(2) + Q.bmm(x.unsqueeze(2)).squeeze(2) + p
        if e > 0:
            rx = rx + A.transpose(1, 2).bmm(y.unsqueeze(2)).squeeze(2)
        rz = G.bmm(x.unsqueeze(2)).squeeze(2) + s - h
        ry = A.bmm(x.unsqueeze(2)).squeeze(2) - b if e > 0 else None
        mu = ((s * z).sum(1) / m).abs()
        pri = rz.norm(2, 1) + (ry.norm(2, 1) if e > 0 else 0.0)
        resids = pri + rx.norm(2, 1) + m * mu
        d = z / s
        try:
            F.factor_kkt(d)
        except RuntimeError:
            break
        if best is None:
            best = dict(r=resids.clone(), x=x.clone(), z=z.clone(), s=s.clone(),
                        y=y.clone() if e > 0 else None)
            nNot = 0
        else:
            I = resids < best["r"]
            nNot = 0 if bool(I.any()) else nNot + 1
            best["r"] = torch.where(I, resids, best["r"])
            Iu = I.unsqueeze(1)
            best["x"] = torch.where(Iu, x, best["x"]); best["z"] = torch.where(Iu, z, best["z"])
            best["s"] = torch.where(Iu, s, best["s"])
            if e > 0:
                best["y"] = torch.where(Iu, y, best["y"])
        if nNot == notImprovedLim or bool(best["r"].max() < eps) or bool(mu.min() > 1e32):
            break
        dx_a, ds_a, dz_a, dy_a = F.solve_kkt(G, A, d, rx, z, rz, ry)
        alpha = torch.min(torch.min(get_step(z, dz_a), get_step(s, ds_a)), torch.ones(B, dtype=Q.dtype))
        a1 = alpha.unsqueeze(1)
        sig = (((s + a1 * ds_a) * (z + a1 * dz_a)).sum(1) / (s * z).sum(1)) ** 3
        rs_c = (-(mu * sig).unsqueeze(1) + ds_a * dz_a) / s
        dx_c, ds_c, dz_c, dy_c = F.solve_kkt(G, A, d, torch.zeros(B, n, dtype=Q.dtype), rs_c, zeros_m,
                                             torch.zeros(B, e, dtype=Q.dtype) if e > 0 else None)
        dx, ds, dz = dx_a + dx_c, ds_a + ds_c, dz_a + dz_c
        alpha = torch.min(0.999 * torch.min(get_step(z, dz), get_step(s, ds)), torch.ones(B, dtype=Q.dtype))
        a1 = alpha.unsqueeze(1)
        x = x + a1 * dx; s = s + a1 * ds; z = z + a1 * dz
        if e > 0:
            y = y + a1 * (dy_a + dy_c)
    return best["x"], best["y"], best["z"], best["s"]


def backward(G, A, F, zhat, lam, slacks, nus, dl):
    """qp.py:128-182 for fully batched inputs."""
    B, m, n = G.shape
    e = F.e
    d = lam.clamp(min=1e-8) / slacks.clamp(min=1e-8)
    F.factor_kkt(d)
    zeros_m = torch.zeros(B, m, dtype=G.dtype)
    dx, _, dlam, dnu = F.solve_kkt(G, A, d, dl, zeros_m, zeros_m,
                                   torch.zeros(B, e, dtype=G.dtype) if e > 0 else None)
    dG = dlam.unsqueeze(2) * zhat.unsqueeze(1) + lam.unsqueeze(2) * dx.unsqueeze(1)
    dQ = 0.5 * (dx.unsqueeze(2) * zhat.unsqueeze(1) + zhat.unsqueeze(2) * dx.unsqueeze(1))
    dA = db = None
    if e > 0:
        dA = dnu.unsqueeze(2) * zhat.unsqueeze(1) + nus.unsqueeze(2) * dx.unsqueeze(1)
        db = -dnu
    return dQ, dx, dG, -dlam, dA, db


def qp_fwd_bwd(Q, p, G, h, A, b, dl, eps=1e-12, notImprovedLim=3, maxIter=20):
    """QPFunction()(Q,p,G,h,A,b) + backward(dl) for fully batched torch CPU tensors."""
    with torch.no_grad():
        torch.linalg.cholesky(Q)                                        # qp.py:81-85 (SPD check)
        F = Factors(Q, G, A)
        x, y, z, s = forward(Q, p, G, h, A, b, F, eps, notImprovedLim, maxIter)
        grads = backward(G, A, F, x, z, s, y, dl)
    return x, z, s, y, grads
