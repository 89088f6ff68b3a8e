"""Quick device timing of setup / forward / backward at one config (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qpth_b200 import QPFunction
from qpth_b200.problems import random_qp_batch
B, n, m, e = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 100, 100, 0))]
pr = random_qp_batch(B, n, m, e, seed=0)
dev = "cuda:0"
t = {k: (torch.tensor(v, dtype=torch.float64, device=dev, requires_grad=True) if v.size else torch.Tensor().to(dev).double()) for k, v in pr.items() if k != "dl"}
dl = torch.ones(B, n, dtype=torch.float64, device=dev)
f = QPFunction(verbose=-1, check_Q_spd=False)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for rep in range(6):
    for v in t.values():
        if v.requires_grad: v.grad = None
    ev[0].record(); z = f(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"]); ev[1].record(); z.backward(dl); ev[2].record()
    torch.cuda.synchronize()
    it = f.last_solve().iters
    print("rep %d fwd %.3f ms bwd %.3f ms  -> %.0f QPs/s  iters min/mean/max %d/%.1f/%d" % (rep, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), B / (ev[0].elapsed_time(ev[2]) * 1e-3), it.min(), it.float().mean(), it.max()))
