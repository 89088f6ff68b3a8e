import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qpth_b200 import QPFunction
from qpth_b200.problems import random_qp_batch
B, n, m, e = [int(x) for x in sys.argv[1:5]]
pr = random_qp_batch(B, n, m, e, seed=0)
dev = "cuda:0"
t = {k: (torch.tensor(v, dtype=torch.float64, device=dev, requires_grad=True) if v.size else torch.Tensor().to(dev).double()) for k, v in pr.items() if k != "dl"}
f = QPFunction(verbose=-1, check_Q_spd=False, maxIter=int(sys.argv[5]) if len(sys.argv) > 5 else 20)
z = f(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"]); z.backward(torch.ones_like(z))
torch.cuda.synchronize()
print("iters", f.last_solve().iters.tolist(), "resid", f.last_solve().best_resid.tolist())
