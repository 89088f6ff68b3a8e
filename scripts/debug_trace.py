import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qpth_b200.qp as qp
qp.TRACE = True
from oracle.cases import load_case
from tests.test_gpu_parity import _run
from qpth_b200 import QPFunction
name, w = sys.argv[1], int(sys.argv[2])
prob, gold, full = load_case(name, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
DEV = "cuda:0"
t = {k: (torch.tensor(np.asarray(prob[k]), dtype=torch.float64, device=DEV) if np.asarray(prob[k]).size else torch.Tensor().to(DEV).double()) for k in ("Q", "p", "G", "h", "A", "b")}
f = QPFunction(verbose=-1)
z = f(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"])
st = f.last_solve()
tr = st.trace[w].cpu().numpy()
for i in range(int(st.iters[w])):
    print(i, "pri %.3e dual %.3e mu %.3e resid %.6e" % tuple(tr[i]))
