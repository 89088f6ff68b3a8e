"""Per-QP parity errors of one golden case on the GPU (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.cases import load_case, proj
from tests.parity import rel_rows
from tests.test_gpu_parity import _run
name = sys.argv[1]
prob, gold, full = load_case(name, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
out = _run(prob)
np.set_printoptions(precision=2, linewidth=200)
print("iters", np.bincount(out["iters"]))
for k in ("zhat", "lam", "slacks"):
    print(k, "max rel", rel_rows(out[k], gold[k]).max())
for k, g in zip(("dQ", "dp", "dG", "dh", "dA", "db"), out["grads"]):
    if g is None: continue
    ref = gold[k] if k in gold else gold[k + "_proj"]
    if k not in gold: g = g @ proj(g.shape[-1])
    e = rel_rows(g, ref, floor=1e-4)
    print(k, "max", e.max(), "median", np.median(e), "argmax", e.argmax())
# separate forward from backward error: oracle backward on the GPU's own forward outputs
from oracle import pdipm_oracle as orc
B = out["zhat"].shape[0]; n = out["zhat"].shape[1]
def bc(k, nd):
    v = np.asarray(prob[k], dtype=float)
    return (np.broadcast_to(v[None], (B,) + v.shape).copy(), True) if v.ndim == nd - 1 else (v, False)
(Q, Qe), (G, Ge) = bc("Q", 3), bc("G", 3)
A = np.zeros((B, 0, n)); 
if np.asarray(prob["A"]).size: A, Ae = bc("A", 3)
F = orc.Factors(Q, G, A)
gr = orc.backward(Q, G, A, F, out["zhat"], out["lam"], out["slacks"], out["nus"], prob["dl"].reshape(B, n), (False,)*6)
for k, g, r in zip(("dQ", "dp", "dG", "dh"), out["grads"], gr):
    if g.shape != r.shape: r = r.mean(0)
    print("GPU bwd vs oracle bwd on GPU duals:", k, rel_rows(g, r, floor=1e-4).max())
w = int(np.argmax(rel_rows(out["grads"][1], gold["dp"], floor=1e-4)))
print("worst QP", w, "iters", out["iters"][w], "best_resid", out["best_resid"][w])
lam_g, lam_r, s_g, s_r = out["lam"][w], gold["lam"][w], out["slacks"][w], gold["slacks"][w]
dg = np.maximum(lam_g, 1e-8) / np.maximum(s_g, 1e-8); dr = np.maximum(lam_r, 1e-8) / np.maximum(s_r, 1e-8)
rel = np.abs(dg - dr) / dr
idx = np.argsort(-rel)[:6]
print("largest d mismatches", [(int(i), "%.2e" % rel[i], "lam %.2e/%.2e s %.2e/%.2e" % (lam_g[i], lam_r[i], s_g[i], s_r[i])) for i in idx])
