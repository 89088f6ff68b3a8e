"""Cycle accounting of k_forward_fast phases (needs libqpth_b200_timing.so built with -DQPB_TIMING)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qpth_b200 import _lib
_lib.LIB_PATH = os.environ.get("QPB200_TIMING_LIB") or os.path.join(os.path.dirname(_lib.LIB_PATH), "libqpth_b200_timing.so")
from qpth_b200 import QPFunction
from qpth_b200.problems import random_qp_batch
B, n, m, e = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 100, 100, 0))]
lib = _lib.load()
lib.qpb200_debug_timing.restype = ctypes.c_int
lib.qpb200_debug_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
pr = random_qp_batch(B, n, m, e, seed=0)
dev = "cuda:0"
t = {k: (torch.tensor(v, dtype=torch.float64, device=dev) if v.size else torch.Tensor().to(dev).double()) for k, v in pr.items() if k != "dl"}
f = QPFunction(verbose=-1, check_Q_spd=False)
f(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"]); torch.cuda.synchronize()
lib.qpb200_debug_timing(None, 1)
f(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"]); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
lib.qpb200_debug_timing(buf, 0)
torch.cuda.synchronize(); it = int(f.last_solve().iters.cpu()[0]); print('iters', f.last_solve().iters.cpu()[:8].tolist(), 'resid', f.last_solve().best_resid.cpu()[:3].tolist())
names = {0: "make_ctx (TMA W,L)", 1: "load vectors", 2: "whiten", 3: "loop misc/update (prev)", 4: "matvec_cols (r~x)", 5: "matvec_rows2", 6: "residual elementwise", 7: "tri_norm2", 8: "reduce_sum4",
         9: "best/exit/aug build", 10: "factor_and_solve tail", 11: "aff step, sigma, rhs", 12: "trsv_fwd (cor)", 13: "trsv_bwd (cor)", 14: "issue_K + combine", 15: "matvec_cols (dx)", 16: "final misc",
         20: "chol: tile load + first factor", 21: "chol: first barrier", 22: "chol: phase A work", 23: "chol: barrier after A", 24: "chol: diag tile k+1 update", 25: "chol: factor8", 26: "chol: (other warps) / end B", 27: "chol: barrier after B",
         40: "[warp1] gap", 41: "[warp1] S_k rows", 42: "[warp1] named barrier wait", 43: "[warp1] U_k update", 44: "[warp1] step barrier wait", 17: "pform: T_k blocks", 18: "vg_sum2: entry gap", 19: "vg_sum2: 2 warp sums", 22: "vg_sum2: named barrier", 23: "vec_resid: f_div", 25: "pform: P conversion loop (rest)", 29: "pform: P conversion, block column 0", 27: "ptrsv_fwd: sweep", 28: "pform: P conversion (+17)", 34: "vec_resid: loop 1", 35: "vec_resid: 2 x sum2", 36: "vec_resid: scalars", 41: "vec_affine: loop 1", 43: "vec_affine: min2", 45: "vec_affine: loop2+sum2+div",
         60: "[probe] barrier+loads", 61: "[probe] 4 warp_sum, 8 warps", 62: "[probe] 4 warp_sum, warps 0-3", 63: "[probe] 4 warp_min, warps 0-3", 77: "[probe] 16 dependent DMMA, 8 warps", 78: "[probe] 16 dependent DMMA, warp 0", 79: "[probe] 16 dependent DFMA", 93: "[probe] closing barrier", 94: "-",
         30: "wait K copy", 31: "diag add + barrier", 32: "chol exit", 33: "trsv_bwd (aff)"}
tot = sum(buf[i] for i in range(40)) + buf[45] + buf[41] + buf[43]
print("QP 0 of block 0: %d iterations; total %d cycles (%.1f us @1.965GHz); per iteration %.0f" % (it, tot, tot / 1965.0, tot / (it + 1)))
for i in list(range(48)) + [60, 61, 62, 63, 77, 78, 79, 93, 94]:
    if buf[i]:
        print("%2d %-34s %9d cyc  %5.1f%%   per-iter %7.0f" % (i, names.get(i, "?"), buf[i], 100.0 * buf[i] / tot, buf[i] / (it + 1)))

nf = it + 1
print("per Cholesky step k (cycles per factorization): S_k[w1]  U_k[w1] | s_k[w0]  F_k+1[w0]  wait[w0]")
for k in range(13):
    print("%2d  %6.0f %6.0f | %6.0f %6.0f %6.0f" % (k, buf[64 + k] / nf, buf[48 + k] / nf, buf[96 + k] / nf, buf[80 + k] / nf, buf[112 + k] / nf))
