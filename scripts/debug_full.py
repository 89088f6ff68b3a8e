import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pdipm_oracle as orc
from qpth_b200.problems import random_qp_batch
from tests.parity import rel_rows
from tests.test_gpu_parity import _run
cfg = dict(nBatch=128, nz=100, nineq=100, neq=0)
pr = random_qp_batch(seed=11, **cfg)
out = _run(pr)
ref = orc.qp_solve(pr["Q"], pr["p"], pr["G"], pr["h"], pr["A"], pr["b"], pr["dl"], per_qp=True)
print("z", rel_rows(out["zhat"], ref["zhat"]).max())
for k,(g, r) in enumerate(zip(out["grads"], ref["grads"])):
    if r is not None: 
        e = rel_rows(g, r, floor=1e-4); print(k, e.max(), e.argmax())
di = out["iters"] - ref["info"]["iters"]
print("iters diff", np.bincount(di - di.min()), di.min(), "gpu", out["iters"][:16], "ref", ref["info"]["iters"][:16])
print("best resid gpu", out["best_resid"][:8], "ref", ref["info"]["best_resids"][:8])
