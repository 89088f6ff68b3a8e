"""`QPFunction` — drop-in for `qpth.qp.QPFunction` backed by the sm_100a kernels.

Mirrors the reference's autograd boundary (`qpth/qp.py:18-183`): same factory
signature, same broadcasting of un-batched parameters (`qpth/util.py:44-59`),
same error strings, same gradient conventions (batch MEAN for un-batched
inputs, `dA = db = None` when there are no equality constraints, symmetrised
dQ).  The solver behind it is `libqpth_b200.so` (include/qpth_b200.h); there is
no CPU fallback — without the library or without a CUDA device the call raises.

Differences from the reference, all documented in DESIGN.md:
  * every QP is solved with the reference's nBatch=1 semantics (the batch-global
    exit tests and get_step fill value are applied per QP);
  * arithmetic is fp64 on the device whatever the input dtype (fp32 inputs are
    promoted and the results cast back);
  * CPU tensors are accepted: they are copied to the current CUDA device,
    solved there, and the results returned on the CPU;
  * `solver=QPSolvers.CVXPY` (qp.py:97-120) needs the `cvxpy` package for its forward; its backward, and the
    stand-alone `QPSolutionFunction` for solutions produced by any other solver, live in `solution.py`.
"""
import ctypes
from enum import Enum

import torch
from torch.autograd import Function

from . import _lib
from .util import check_shapes, expandParam, extract_nBatch

INACC_ERR = """
--------
qpth warning: Returning an inaccurate and potentially incorrect solution.

Some residual is large.
Your problem may be infeasible or difficult.

You can try using the CVXPY solver to see if your problem is feasible
and you can use the verbose option to check the convergence status of
our solver while increasing the number of iterations.

Advanced users:
You can also try to enable iterative refinement in the solver:
https://github.com/locuslab/qpth/issues/6
--------
"""


# The reference stops a batch when NO QP improved for `notImprovedLim` consecutive iterations
# (batch.py:127-143).  Applied per QP that rule would abandon QPs that merely pause while still far from
# a solution (resids is not monotone early on); it is therefore only applied once a QP's best residual
# is below STALL_TOL, i.e. when it is stalling at its rounding floor like the rest of its batch would.
STALL_TOL = 1e-6
# The reference returns the argmin-resids iterate (batch.py:126-139).  At the rounding floor successive
# iterates tie to within noise while mu still shrinks ~1000x per iteration, and which one wins decides
# whether the backward pass's 1e-8 clamps (qp.py:148) are saturated.  Among iterates within BEST_TIE of
# the minimum the latest is returned (1.0 restores the literal rule).
BEST_TIE = 1.5
TRACE = False     # keep per-iteration residual traces on st.trace even when verbose != 1 (diagnostics)
# Kernel choice where a shape has several product-form variants (plan.pf2_ok / plan.pf3_ok, e.g. nz = nineq = 100):
#   "latency"    one QP per SM, W / chol(Q) / factor in shared memory: the shortest time for ONE batch <= #SMs;
#   "throughput" three (else two) QPs per SM, W and chol(Q) read from L2: +34 % QPs/s once the GPU is full (a large
#                batch, or several batches in flight on several streams), 1.5x the latency of a lone 128-QP batch
#                (profiles/r2z_kernel_times.txt);
#   "auto"       throughput when the batch alone exceeds the SM count, else latency.
# Set qpth_b200.qp.MODE (or QPTH_B200_MODE) before the call. The variants differ only in the summation order of the
# W / chol(Q) passes: results agree to ~1e-12 relative (tests/test_gpu_parity.py pins 1e-10), not bit for bit.
import os as _os
MODE = _os.environ.get("QPTH_B200_MODE", "auto")
# The reference's DEFAULT options (check_Q_spd=True, verbose=0) make every forward read two flags back from the device
# before it returns ('Q is not SPD.' qp.py:81-85; the inaccurate-solution banner batch.py:205-206): one blocking host
# read per call, which caps a training loop at ~72k QPs/s on a B200 (profiles/r2z_bench.json, e2e.default_options)
# where the asynchronous options reach 230k. LAZY_CHECKS = True keeps the diagnostics but DEFERS them: the flags are
# copied to pinned host memory asynchronously and examined at the next QPFunction call, in backward, or by
# flush_checks() - the error / banner then refers to an EARLIER call. Off by default: the reference raises at once.
LAZY_CHECKS = _os.environ.get("QPTH_B200_LAZY_CHECKS", "0") == "1"
_pending = []
import threading as _threading
_pending_lock = _threading.Lock()      # forward runs on the caller's thread, backward on the autograd engine's


def flush_checks(wait=True):
    """Examine the deferred diagnostics (LAZY_CHECKS) of earlier forward calls; wait=False only those already on the host."""
    while True:
        with _pending_lock:
            if not _pending:
                return
            ev, host, chk, banner = _pending[0]
            if not wait and not ev.query():
                return
            _pending.pop(0)
        ev.synchronize()
        bad_spd, inacc = host.tolist()
        if banner and inacc:
            print(INACC_ERR)
        if chk and bad_spd:
            raise RuntimeError('Q is not SPD. (reported by a deferred check, qpth_b200.qp.LAZY_CHECKS)')


class QPSolvers(Enum):
    PDIPM_BATCHED = 1
    CVXPY = 2


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev64(t, device):
    """fp64, contiguous, on `device`; un-batched tensors stay un-batched (stride 0 is passed to C)."""
    return t.detach().to(device=device, dtype=torch.float64).contiguous()


class _Solved:
    """State carried from forward to backward (the reference's ctx.Q_LU / S_LU / R / nus / lams / slacks)."""
    __slots__ = ("plan", "nBatch", "nsys", "L", "W", "K", "zhat", "lam", "slacks", "nus",
                 "iters", "best_resid", "scratch", "device", "trace")


def solve_forward(Q_, p_, G_, h_, A_, b_, eps=1e-12, verbose=0, notImprovedLim=3, maxIter=20,
                  check_Q_spd=True):
    """pre_factor_kkt + forward on the device. Inputs follow QPFunction's conventions. Returns _Solved."""
    # rank errors exactly as expandParam raises them (util.py:44-50), then every trailing dimension / batch size:
    # pure host logic, done before anything touches the device
    nBatch, nz, nineq, neq = check_shapes(Q_, p_, G_, h_, A_, b_)
    if _pending:
        flush_checks(wait=False)
    assert neq > 0 or nineq > 0                         # qp.py:89
    if nineq == 0:
        raise RuntimeError('qpth_b200: nineq == 0 is not supported (the reference unpacks G.size() at qp.py:87)')
    assert maxIter >= 1
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.QpthB200Error("qpth_b200: no CUDA device available (there is no CPU fallback).")
    device = Q_.device if Q_.is_cuda else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(device):
        Q, p, G, h = (_dev64(t, device) for t in (Q_, p_, G_, h_))
        A = _dev64(A_, device) if neq > 0 else None
        b = _dev64(b_, device) if neq > 0 else None
        # one QP per SM (W, chol(Q) and the factor all in shared memory) has the lower latency; two QPs per SM (W and
        # chol(Q) read from L2) the higher throughput once more QPs are in flight than the GPU has SMs (MODE above)
        two = MODE == "throughput" or (MODE == "auto" and nBatch > _lib.sm_count(device.index or 0))
        plan = _lib.plan_for(nz, nineq, neq, two=two)

        def stride(t, nd, per):
            return per if (t is not None and t.dim() == nd) else 0

        sQ, sG = stride(Q, 3, nz * nz), stride(G, 3, nineq * nz)
        sA = stride(A, 3, neq * nz)
        sp, sh, sb = stride(p, 2, nz), stride(h, 2, nineq), stride(b, 2, neq)
        nsys = nBatch if (sQ or sG or sA) else 1
        st = _Solved()
        st.plan, st.nBatch, st.nsys, st.device = plan, nBatch, nsys, device
        f64 = dict(dtype=torch.float64, device=device)
        st.L = torch.empty(nsys * plan.L_elems, **f64)
        st.W = torch.empty(nsys * plan.W_elems, **f64)
        st.K = torch.empty(nsys * plan.K_elems, **f64)
        spd = torch.zeros(nsys, dtype=torch.int32, device=device)
        nscr = max(nsys * plan.setup_scratch_elems, nBatch * plan.solve_scratch_elems)
        st.scratch = torch.empty(nscr, **f64) if nscr > 0 else None
        _lib.check(lib.qpb200_pre_factor_kkt(
            ctypes.byref(plan), nsys, _ptr(Q), sQ, _ptr(G), sG, _ptr(A), sA,
            _ptr(st.L), _ptr(st.W), _ptr(st.K), _ptr(spd), _ptr(st.scratch), _stream()))
        st.zhat = torch.empty(nBatch, nz, **f64)
        st.lam = torch.empty(nBatch, nineq, **f64)
        st.slacks = torch.empty(nBatch, nineq, **f64)
        st.nus = torch.empty(nBatch, neq, **f64) if neq > 0 else None
        st.iters = torch.empty(nBatch, dtype=torch.int32, device=device)
        st.best_resid = torch.empty(nBatch, **f64)
        st.trace = torch.full((nBatch, int(maxIter), 4), float('nan'), **f64) if (verbose == 1 or TRACE) else None
        _lib.check(lib.qpb200_forward(
            ctypes.byref(plan), nBatch, _ptr(p), sp, _ptr(h), sh, _ptr(b), sb,
            _ptr(st.L), _ptr(st.W), _ptr(st.K), 1 if nsys > 1 else 0,
            float(eps), float(STALL_TOL), float(BEST_TIE), int(notImprovedLim), int(maxIter),
            _ptr(st.zhat), _ptr(st.lam), _ptr(st.slacks), _ptr(st.nus),
            _ptr(st.iters), _ptr(st.best_resid), _ptr(st.trace), _ptr(st.scratch), _stream()))
        # One host read for both diagnostics (the reference syncs many times per iteration):
        # 'Q is not SPD.' (qp.py:81-85) and the inaccurate-solution banner, printed iff
        # best resids max > 1 and verbose >= 0 (batch.py:141-142,205-206).
        if check_Q_spd or verbose >= 0:
            flags = torch.stack([spd.any(), (st.best_resid.max() > 1.)])
            if LAZY_CHECKS:
                # the two flags travel to a pinned host buffer asynchronously; they are examined (and 'Q is not SPD.' raised,
                # the banner printed) at the next QPFunction call, in this call's backward, or by flush_checks()
                host = torch.empty(2, dtype=flags.dtype).pin_memory()
                host.copy_(flags, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                _pending.append((ev, host, bool(check_Q_spd), verbose >= 0))
            else:
                bad_spd, inacc = flags.tolist()
                if check_Q_spd and bad_spd:
                    raise RuntimeError('Q is not SPD.')
                if verbose >= 0 and inacc:
                    print(INACC_ERR)
        if verbose == 1:
            # batch.py:115-117: per-iteration batch means; a QP that has already stopped contributes
            # the values of its last iteration (in the reference every QP runs every iteration)
            tr = st.trace.cpu()
            for i in range(int(st.iters.max())):
                row = tr[:, i, :]
                last = tr[torch.arange(nBatch), (st.iters.cpu().long() - 1).clamp(min=0), :]
                row = torch.where(torch.isnan(row), last, row)
                print('iter: {}, pri_resid: {:.5e}, dual_resid: {:.5e}, mu: {:.5e}'.format(
                    i, row[:, 0].mean(), row[:, 1].mean(), row[:, 2].mean()))
    return st


def solve_backward(st, dl_dzhat, mean_flags, want):
    """QPFunctionFn.backward on the device. mean_flags / want: 6-tuples for (Q,p,G,h,A,b)."""
    if _pending:
        flush_checks(wait=False)
    lib = _lib.load()
    plan, B, device = st.plan, st.nBatch, st.device
    nz, nineq, neq = plan.nz, plan.nineq, plan.neq
    f64 = dict(dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        dl = dl_dzhat.detach().to(device=device, dtype=torch.float64).contiguous().view(B, nz)
        shapes = [(nz, nz), (nz,), (nineq, nz), (nineq,), (neq, nz), (neq,)]
        outs = []
        for k in range(6):
            if not want[k] or (k >= 4 and neq == 0):
                outs.append(None)
            else:
                shp = shapes[k] if mean_flags[k] else (B,) + shapes[k]
                outs.append(torch.empty(*shp, **f64))
        dxv = torch.empty(B, nz, **f64)
        dlamv = torch.empty(B, nineq, **f64)
        dnuv = torch.empty(B, neq, **f64) if neq > 0 else None
        args = []
        for k in range(6):
            args += [_ptr(outs[k]), 1 if mean_flags[k] else 0]
        _lib.check(lib.qpb200_backward(
            ctypes.byref(plan), B, _ptr(dl), _ptr(st.zhat), _ptr(st.lam), _ptr(st.slacks), _ptr(st.nus),
            _ptr(st.L), _ptr(st.W), _ptr(st.K), 1 if st.nsys > 1 else 0,
            *args, _ptr(dxv), _ptr(dlamv), _ptr(dnuv), _ptr(st.scratch), _stream()))
    return outs


def QPFunction(eps=1e-12, verbose=0, notImprovedLim=3, maxIter=20, solver=QPSolvers.PDIPM_BATCHED,
               check_Q_spd=True):
    """Factory with the reference's signature (`qpth/qp.py:18-20`); returns `Function.apply`."""
    if solver == QPSolvers.CVXPY:
        # qp.py:97-120,142-143: per-sample CVXPY solve on the CPU, then pre_factor_kkt + the same backward.
        from .solution import QPSolutionFunction, cvxpy_forward

        def apply_cvxpy(Q_, p_, G_, h_, A_, b_):
            nBatch = extract_nBatch(Q_, p_, G_, h_, A_, b_)
            Q, p, G, h, A, b = (expandParam(X, nBatch, nd)[0]
                                for X, nd in ((Q_, 3), (p_, 2), (G_, 3), (h_, 2), (A_, 3), (b_, 2)))
            zhats, nus, lams, slacks = cvxpy_forward(Q, p, G, h, A, b)
            return QPSolutionFunction(check_Q_spd)(Q_, p_, G_, h_, A_, b_, zhats, lams, slacks, nus)

        return apply_cvxpy
    if solver != QPSolvers.PDIPM_BATCHED:
        assert False                                     # qp.py:121-122

    _last = [None]

    class QPFunctionFn(Function):
        @staticmethod
        def forward(ctx, Q_, p_, G_, h_, A_, b_):
            """Solve a batch of QPs  argmin_z 1/2 z^T Q z + p^T z  s.t. Gz <= h, Az = b.

            Q (nBatch,nz,nz)|(nz,nz); p (nBatch,nz)|(nz); G (nBatch,nineq,nz)|(nineq,nz);
            h (nBatch,nineq)|(nineq); A (nBatch,neq,nz)|(neq,nz)|empty; b (nBatch,neq)|(neq)|empty.
            Returns zhat (nBatch, nz).  (qp.py:23-125)
            """
            st = solve_forward(Q_, p_, G_, h_, A_, b_, eps, verbose, notImprovedLim, maxIter, check_Q_spd)
            ctx.st = st
            _last[0] = st
            ctx.neq, ctx.nineq, ctx.nz = st.plan.neq, st.plan.nineq, st.plan.nz
            zhats = st.zhat.to(device=Q_.device, dtype=Q_.dtype)
            ctx.save_for_backward(zhats, Q_, p_, G_, h_, A_, b_)
            # parity with the reference's ctx attributes (device fp64 views)
            ctx.lams, ctx.slacks, ctx.nus = st.lam, st.slacks, st.nus
            return zhats

        @staticmethod
        def backward(ctx, dl_dzhat):
            zhats, Q, p, G, h, A, b = ctx.saved_tensors
            nBatch = extract_nBatch(Q, p, G, h, A, b)
            flags = [expandParam(X, nBatch, nd)[1]
                     for X, nd in ((Q, 3), (p, 2), (G, 3), (h, 2), (A, 3), (b, 2))]   # qp.py:131-136
            want = list(ctx.needs_input_grad)
            outs = solve_backward(ctx.st, dl_dzhat, flags, want)
            grads = []
            for X, g in zip((Q, p, G, h, A, b), outs):
                grads.append(None if g is None else g.to(device=X.device, dtype=X.dtype))
            return tuple(grads)

    def apply(Q_, p_, G_, h_, A_, b_):
        if G_.nelement() == 0 and h_.nelement() == 0 and A_.nelement() > 0:
            # equality-constrained QP: an extension (the reference cannot run nineq == 0); one KKT solve, eqonly.py
            from .eqonly import solve_equality_qp
            return solve_equality_qp(Q_, p_, A_, b_, check_Q_spd)
        return QPFunctionFn.apply(Q_, p_, G_, h_, A_, b_)

    # diagnostics the reference keeps on ctx (nus / lams / slacks) plus per-QP iteration counts
    apply.last_solve = lambda: _last[0]
    return apply
