"""Equality-constrained QPs (nineq == 0, neq > 0) behind `QPFunction`: an EXTENSION of the reference's surface.

The reference cannot run this case - it unpacks `G.size()` (`qpth/qp.py:87`) and takes the minimum of an empty slack
vector (`solvers/pdipm/batch.py:77`) - although its own algebra covers it: without inequalities the optimum is ONE KKT
solve, the one `forward` performs for its initial point (`batch.py:61-67`: `solve_kkt(p, 0, -h, -b)` with d = 1), and
the backward pass is the same solve with `dl_dzhat` on the right-hand side (`qp.py:148-177`). Here both are the
stand-alone `pre_factor_kkt` + `solve_kkt` kernels (`qpth_b200/kkt.py`, C ABI `qpb200_pre_factor_kkt_reg` /
`qpb200_solve_kkt_reg`) with one decoupled dummy inequality row (G = 0, h = 1, d = 1: the reduced matrix gets a unit
diagonal entry for it and nothing else), so no interior-point iteration - and no step length over a slack that never
moves - is involved. Gradient conventions are the reference's: symmetrised dQ, `.mean(0)` for un-batched inputs.
"""
import torch
from torch.autograd import Function

from . import _lib, kkt
from .util import bger, check_shapes, expandParam

_factor = kkt._Factored          # (tests inject a CPU stand-in here)


def _target_device(Q_):
    if not torch.cuda.is_available():
        raise _lib.QpthB200Error("qpth_b200: no CUDA device available (there is no CPU fallback).")
    return Q_.device if Q_.is_cuda else torch.device("cuda", torch.cuda.current_device())


class QPEqualityFn(Function):
    @staticmethod
    def forward(ctx, Q_, p_, A_, b_, check_Q_spd):
        empty = Q_.new_empty(0)
        nBatch, nz, nineq, neq = check_shapes(Q_, p_, empty, empty, A_, b_)
        assert neq > 0 or nineq > 0                         # qp.py:89
        device = _target_device(Q_)
        f64 = dict(dtype=torch.float64, device=device)
        batched = []
        flags = []
        for X, nd in ((Q_, 3), (p_, 2), (A_, 3), (b_, 2)):
            Xe, was_unbatched = expandParam(X.detach().to(**f64), nBatch, nd)
            batched.append(Xe.contiguous())
            flags.append(was_unbatched)
        Q, p, A, b = batched
        F = _factor(Q, torch.zeros(nBatch, 1, nz, **f64), A, 0.0)          # one dummy row: G = 0 (h = 1, d = 1 below)
        if check_Q_spd and bool(F.spd.any()):
            raise RuntimeError('Q is not SPD.')
        one, zero = torch.ones(nBatch, 1, **f64), torch.zeros(nBatch, 1, **f64)
        zhat, _, _, nus = F.solve(one, p, zero, -one, -b)                   # K [x s z y] = -[p 0 -h -b]
        ctx.F, ctx.flags, ctx.one, ctx.zero = F, flags, one, zero
        ctx.zhat64, ctx.nus = zhat, nus
        ctx.meta = [(X.device, X.dtype) for X in (Q_, p_, A_, b_)]
        return zhat.to(device=Q_.device, dtype=Q_.dtype)

    @staticmethod
    def backward(ctx, dl_dzhat):
        z, nus = ctx.zhat64, ctx.nus
        dl = dl_dzhat.detach().to(device=z.device, dtype=torch.float64).contiguous().view_as(z)
        dx, _, _, dnu = ctx.F.solve(ctx.one, dl, ctx.zero, ctx.zero, torch.zeros_like(nus))
        grads = [0.5 * (bger(dx, z) + bger(z, dx)),         # qp.py:175-177
                 dx,                                        # qp.py:150
                 bger(dnu, z) + bger(nus, dx),              # qp.py:166-167
                 -dnu]                                      # qp.py:168
        out = []
        for g, unb, (dev, dt), need in zip(grads, ctx.flags, ctx.meta, ctx.needs_input_grad[:4]):
            if not need:
                out.append(None)
                continue
            out.append((g.mean(0) if unb else g).to(device=dev, dtype=dt))
        return tuple(out) + (None,)


def solve_equality_qp(Q_, p_, A_, b_, check_Q_spd=True):
    """z* of  argmin 1/2 z'Qz + p'z  s.t. Az = b  (differentiable in Q, p, A, b)."""
    return QPEqualityFn.apply(Q_, p_, A_, b_, check_Q_spd)
