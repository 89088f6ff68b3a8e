import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qpth_b200 import _lib
_lib.LIB_PATH = os.environ.get("QPB200_TIMING_LIB") or os.path.join(os.path.dirname(_lib.LIB_PATH), "libqpth_b200_timing.so")
from qpth_b200.problems import random_qp_batch
B, n, m, e = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 100, 100, 0))]
lib = _lib.load(); lib.qpb200_debug_timing.restype = ctypes.c_int; lib.qpb200_debug_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
plan = _lib.plan_for(n, m, e); pr = random_qp_batch(B, n, m, e, seed=0); dev = "cuda:0"
tt = lambda a: torch.tensor(a, dtype=torch.float64, device=dev).contiguous()
Q, G, A = tt(pr["Q"]), tt(pr["G"]), tt(pr["A"])
f64 = dict(dtype=torch.float64, device=dev)
L = torch.empty(B * plan.L_elems, **f64); W = torch.empty(B * plan.W_elems, **f64); K = torch.empty(B * plan.K_elems, **f64)
spd = torch.zeros(B, dtype=torch.int32, device=dev)
P = lambda x: ctypes.c_void_p(x.data_ptr()); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(2):
    lib.qpb200_debug_timing(None, 1)
    _lib.check(lib.qpb200_pre_factor_kkt(ctypes.byref(plan), B, P(Q), n * n, P(G), m * n, P(A) if e else None, e * n, P(L), P(W), P(K), P(spd), None, st))
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)(); lib.qpb200_debug_timing(buf, 0)
names = {34: "stage Q,G", 35: "chol(Q)", 36: "T blocks", 37: "W = rows L^-T", 38: "write L, W", 39: "K = W W^T", 45: "eq partial chol", 46: "write K"}
print("plan: setup_fast=%d setup_pf=%d" % (plan.setup_fast, plan.setup_pf))
for i in (34, 35, 36, 37, 38, 39, 45, 46): print("%-16s %8d cycles" % (names[i], buf[i]))
print("chol(Q) per step (chain warp): publish / F_k+1 / wait:", [(buf[96 + k], buf[80 + k], buf[112 + k]) for k in range(13)])
print("warp 1: S_k, U_k per step:", [(buf[64 + k], buf[48 + k]) for k in range(13)], "rhs rows", buf[45], "panel wait", buf[43], "T wait", buf[42], "step barrier", buf[44])
