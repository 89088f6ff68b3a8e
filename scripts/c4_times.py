"""Where the C4 step (cls-layer pattern: shared Q, G, h; batched p) spends its time: CUDA events around the three C-ABI
calls with ONE shared system (nsys = 1), then the same step through QPFunction (adds the batch-mean kernels)."""
import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qpth_b200 import _lib, QPFunction
from qpth_b200.problems import cls_layer_problem
B, n, m = 64, 200, 200
lib = _lib.load(); plan = _lib.plan_for(n, m, 0, two=(None if os.environ.get('QPB_KT_TWO') is None else os.environ['QPB_KT_TWO'] == '1'))
pr = cls_layer_problem(B, n, m, seed=0); dev = "cuda:0"
tt = lambda a: torch.tensor(a, dtype=torch.float64, device=dev).contiguous()
Q, p, G, h = (tt(pr[k]) for k in ("Q", "p", "G", "h"))
f64 = dict(dtype=torch.float64, device=dev)
L = torch.empty(plan.L_elems, **f64); W = torch.empty(plan.W_elems, **f64); K = torch.empty(plan.K_elems, **f64)
spd = torch.zeros(1, dtype=torch.int32, device=dev)
z = torch.empty(B, n, **f64); lam = torch.empty(B, m, **f64); s = torch.empty(B, m, **f64)
it = torch.empty(B, dtype=torch.int32, device=dev); rr = torch.empty(B, **f64)
dl = torch.ones(B, n, **f64)
gQ, gp, gG, gh = torch.empty(n, n, **f64), torch.empty(B, n, **f64), torch.empty(m, n, **f64), torch.empty(m, **f64)
wx, wl = torch.empty(B, n, **f64), torch.empty(B, m, **f64)
P = lambda x: ctypes.c_void_p(x.data_ptr()) if x is not None else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
scr = torch.empty(max(1, max(plan.setup_scratch_elems, B * plan.solve_scratch_elems)), **f64)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
res = []
for rep in range(5):
    ev[0].record()
    _lib.check(lib.qpb200_pre_factor_kkt(ctypes.byref(plan), 1, P(Q), 0, P(G), 0, None, 0, P(L), P(W), P(K), P(spd), P(scr), st))
    ev[1].record()
    _lib.check(lib.qpb200_forward(ctypes.byref(plan), B, P(p), n, P(h), 0, None, 0, P(L), P(W), P(K), 0, 1e-12, 1e-6, 1.5, 3, 20, P(z), P(lam), P(s), None, P(it), P(rr), None, P(scr), st))
    ev[2].record()
    _lib.check(lib.qpb200_backward(ctypes.byref(plan), B, P(dl), P(z), P(lam), P(s), None, P(L), P(W), P(K), 0, P(gQ), 1, P(gp), 0, P(gG), 1, P(gh), 1, None, 0, None, 0, P(wx), P(wl), None, P(scr), st))
    ev[3].record(); torch.cuda.synchronize()
    res.append([ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(3)])
r = np.array(res[1:]).mean(0)
print("C4 shared system, C ABI: setup %.1f us, forward %.1f us, backward(+mean) %.1f us -> %.0f QPs/s; iters mean %.1f max %d" % (r[0], r[1], r[2], B / (r.sum() * 1e-6), it.float().mean(), it.max()))
t = {k: torch.tensor(pr[k], dtype=torch.float64, device=dev, requires_grad=True) for k in ("Q", "p", "G", "h")}
e = torch.Tensor().to(dev).double()
f = QPFunction(verbose=-1, check_Q_spd=False)
for rep in range(4):
    for v in t.values(): v.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    zz = f(t["Q"], t["p"], t["G"], t["h"], e, e); torch.cuda.synchronize(); t1 = time.perf_counter()
    zz.backward(dl); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("QPFunction: forward %.2f ms, backward %.2f ms (wall, synchronised)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
