"""Kernel-only durations (CUDA events around each C-ABI call) at one config (dev tool)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qpth_b200 import _lib
from qpth_b200.problems import random_qp_batch
B, n, m, e = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 100, 100, 0))]
lib = _lib.load(); plan = _lib.plan_for(n, m, e, two=(None if os.environ.get('QPB_KT_TWO') is None else os.environ['QPB_KT_TWO'] == '1'))
pr = random_qp_batch(B, n, m, e, seed=0); dev = "cuda:0"
tt = lambda a: torch.tensor(a, dtype=torch.float64, device=dev).contiguous()
Q, p, G, h, A, b = (tt(pr[k]) for k in ("Q", "p", "G", "h", "A", "b"))
f64 = dict(dtype=torch.float64, device=dev)
L = torch.empty(B * plan.L_elems, **f64); W = torch.empty(B * plan.W_elems, **f64); K = torch.empty(B * plan.K_elems, **f64)
spd = torch.zeros(B, dtype=torch.int32, device=dev)
z = torch.empty(B, n, **f64); lam = torch.empty(B, m, **f64); s = torch.empty(B, m, **f64); nu = torch.empty(B, max(e, 1), **f64)
it = torch.empty(B, dtype=torch.int32, device=dev); rr = torch.empty(B, **f64)
dl = torch.ones(B, n, **f64)
g = [torch.empty(B, n, n, **f64), torch.empty(B, n, **f64), torch.empty(B, m, n, **f64), torch.empty(B, m, **f64), torch.empty(B, max(e,1), n, **f64), torch.empty(B, max(e,1), **f64)]
wx, wl, wn = torch.empty(B, n, **f64), torch.empty(B, m, **f64), torch.empty(B, max(e, 1), **f64)
P = lambda x: ctypes.c_void_p(x.data_ptr()) if x is not None else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
scr = torch.empty(max(1, B * max(plan.setup_scratch_elems, plan.solve_scratch_elems)), **f64)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
res = []
for rep in range(6):
    ev[0].record()
    _lib.check(lib.qpb200_pre_factor_kkt(ctypes.byref(plan), B, P(Q), n * n, P(G), m * n, P(A) if e else None, e * n, P(L), P(W), P(K), P(spd), P(scr), st))
    ev[1].record()
    _lib.check(lib.qpb200_forward(ctypes.byref(plan), B, P(p), n, P(h), m, P(b) if e else None, e, P(L), P(W), P(K), 1, 1e-12, 1e-6, 1.5, 3, 20, P(z), P(lam), P(s), P(nu) if e else None, P(it), P(rr), None, P(scr), st))
    ev[2].record()
    _lib.check(lib.qpb200_backward(ctypes.byref(plan), B, P(dl), P(z), P(lam), P(s), P(nu) if e else None, P(L), P(W), P(K), 1, P(g[0]), 0, P(g[1]), 0, P(g[2]), 0, P(g[3]), 0, P(g[4]) if e else None, 0, P(g[5]) if e else None, 0, P(wx), P(wl), P(wn) if e else None, P(scr), st))
    ev[3].record(); torch.cuda.synchronize()
    res.append([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(3)])
r = np.array(res[2:]).mean(0)
print("B=%d n=%d m=%d e=%d fast=%d: setup %.1f us, forward %.1f us, backward %.1f us -> %.0f QPs/s (kernel only); iters mean %.1f max %d" % (B, n, m, e, plan.fast, r[0], r[1], r[2], B / (r.sum() * 1e-6), it.float().mean(), it.max()))
