"""Is the e2e leg PCIe-bound or host-bound? (1) host time to ISSUE one e2e step (no sync) vs its device time; (2) the same
step captured in a CUDA graph (H2D copies from pinned memory, QPFunction forward + backward through autograd, D2H copies)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qpth_b200 import QPFunction, qp as qpmod
from qpth_b200.problems import random_qp_batch
qpmod.MODE = "throughput"
dev = torch.device("cuda:0"); B, n, m = 128, 100, 100
NS = int(os.environ.get("NS", "4"))
f = QPFunction(verbose=-1, check_Q_spd=False)
e = torch.Tensor().to(dev).double(); dl = torch.ones(B, n, dtype=torch.float64, device=dev)
hb, dbuf, hout, streams = [], [], [], []
for j in range(NS):
    pr = random_qp_batch(seed=j, nBatch=B, nz=n, nineq=m, neq=0)
    hb.append({k: torch.from_numpy(np.ascontiguousarray(pr[k])).pin_memory() for k in ("Q", "p", "G", "h")})
    dbuf.append({k: torch.empty(v.shape, dtype=torch.float64, device=dev).requires_grad_(True) for k, v in hb[j].items()})
    hout.append({k: torch.empty(s, dtype=torch.float64).pin_memory() for k, s in (("z", (B, n)), ("dQ", (B, n, n)), ("dp", (B, n)), ("dG", (B, m, n)), ("dh", (B, m)))})
    streams.append(torch.cuda.Stream(device=dev))
def body(j):
    src, t, out = hb[j], dbuf[j], hout[j]
    with torch.no_grad():
        for k, v in src.items(): t[k].copy_(v, non_blocking=True)
    for v in t.values(): v.grad = None
    z = f(t["Q"], t["p"], t["G"], t["h"], e, e); z.backward(dl)
    out["z"].copy_(z.detach(), non_blocking=True)
    for k, g in (("dQ", "Q"), ("dp", "p"), ("dG", "G"), ("dh", "h")): out[k].copy_(t[g].grad, non_blocking=True)
def eager(i):
    j = i % NS
    with torch.cuda.stream(streams[j]): body(j)
for i in range(40): eager(i)
torch.cuda.synchronize()
K = 40
t0 = time.perf_counter()
for i in range(K): eager(i)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("eager: host issue %.3f ms/step, wall incl. drain %.3f ms/step -> %.0f QPs/s" % ((t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3, B * K / (t2 - t0)))
# graphs
graphs = []
for j in range(NS):
    s = streams[j]
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): body(j)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        body(j)
    graphs.append(g)
torch.cuda.synchronize()
def graphed(i):
    j = i % NS
    with torch.cuda.stream(streams[j]): graphs[j].replay()
for i in range(20): graphed(i)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(K): graphed(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("graph: host issue %.3f ms/step, wall %.3f ms/step -> %.0f QPs/s" % ((t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3, B * K / (t2 - t0)))
# correctness of the graphed path: outputs equal the eager ones
ref = {k: v.clone() for k, v in hout[0].items()}
eager(0); torch.cuda.synchronize()
print("graph == eager:", all(torch.equal(ref[k], hout[0][k]) for k in ref))
