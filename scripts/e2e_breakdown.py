import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qpth_b200 import QPFunction
from qpth_b200.problems import random_qp_batch
B, n, m = 128, 100, 100
dev = torch.device("cuda:0")
pr = random_qp_batch(B, n, m, 0, seed=0)
hb = {k: torch.from_numpy(np.ascontiguousarray(pr[k])).pin_memory() for k in ("Q", "p", "G", "h")}
host_out = {k: torch.empty(s, dtype=torch.float64).pin_memory() for k, s in (("z", (B, n)), ("dQ", (B, n, n)), ("dp", (B, n)), ("dG", (B, m, n)), ("dh", (B, m)))}
f = QPFunction(verbose=-1, check_Q_spd=False)
e = torch.Tensor().to(dev).double()
dl = torch.ones(B, n, dtype=torch.float64, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
def step(record=False):
    if record: ev[0].record()
    t = {k: v.to(dev, non_blocking=True).requires_grad_(True) for k, v in hb.items()}
    if record: ev[1].record()
    z = f(t["Q"], t["p"], t["G"], t["h"], e, e)
    if record: ev[2].record()
    z.backward(dl)
    if record: ev[3].record()
    host_out["z"].copy_(z.detach(), non_blocking=True)
    for k, g in (("dQ", "Q"), ("dp", "p"), ("dG", "G"), ("dh", "h")):
        host_out[k].copy_(t[g].grad, non_blocking=True)
    if record: ev[4].record()
for i in range(5): step()
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter(); step(True); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("GPU: H2D %.3f fwd %.3f bwd %.3f D2H %.3f ms | host enqueue %.3f ms, total wall %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[3].elapsed_time(ev[4]), (t1 - t0) * 1e3, (t2 - t0) * 1e3))
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): step()
torch.cuda.synchronize(); print("no-sync loop: %.3f ms/step" % ((time.perf_counter() - t0) / 20 * 1e3))
# host-only cost of the API calls (GPU idle time excluded): time the enqueue of fwd+bwd with resident inputs
t = {k: v.to(dev).requires_grad_(True) for k, v in hb.items()}
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(50):
    for v in t.values(): v.grad = None
    z = f(t["Q"], t["p"], t["G"], t["h"], e, e); z.backward(dl)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("resident: host enqueue %.3f ms/step, wall %.3f ms/step" % ((t1 - t0) / 50 * 1e3, (t2 - t0) / 50 * 1e3))

# the same loop but with pre-allocated device input buffers (no allocation per step)
dbuf = {k: torch.empty_like(v, device=dev).requires_grad_(True) for k, v in hb.items()}
def step2():
    with torch.no_grad():
        for k, v in hb.items(): dbuf[k].copy_(v, non_blocking=True)
    for v in dbuf.values(): v.grad = None
    z = f(dbuf["Q"], dbuf["p"], dbuf["G"], dbuf["h"], e, e); z.backward(dl)
    host_out["z"].copy_(z.detach(), non_blocking=True)
    for k, g in (("dQ", "Q"), ("dp", "p"), ("dG", "G"), ("dh", "h")):
        host_out[k].copy_(dbuf[g].grad, non_blocking=True)
for i in range(5): step2()
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(30): step2()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("preallocated inputs, no-sync loop: enqueue %.3f wall %.3f ms/step" % ((t1 - t0) / 30 * 1e3, (t2 - t0) / 30 * 1e3))
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_reserved() / 1e6)
