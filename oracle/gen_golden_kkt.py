"""Golden vectors for the stand-alone KKT solvers of the reference: runs the REAL
`factor_solve_kkt` (batch.py:313-346) and `solve_kkt_ir` (batch.py:244-271) of /root/reference on seeded problems and
writes tests/golden/kkt_*.npz (inputs + the reference's outputs). TEST INFRASTRUCTURE ONLY; build container only.

  python oracle/gen_golden_kkt.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_runner  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def problem(seed, B, nz, nineq, neq, singular=False):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    u = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)
    if singular:                       # PSD Q of rank nz - 3, A with a repeated row (rank neq - 1)
        M = r(B, nz, nz - 3)
        Q = M @ M.transpose(1, 2)
        A = r(B, neq, nz)
        A[:, -1] = A[:, 0]
    else:                              # test.py:190-213 recipe (Q = M M^T), batched
        M = r(B, nz, nz)
        Q = M @ M.transpose(1, 2)
        A = r(B, neq, nz)
    G = r(B, nineq, nz)
    d = u(B, nineq) + (1e-3 if singular else 0.0)
    return dict(Q=Q, G=G, A=A, d=d, rx=u(B, nz), rs=u(B, nineq), rz=u(B, nineq), ry=u(B, neq))


CASES = {"kkt_small": (1, 2, 5, 4, 3, False), "kkt_c3": (2, 8, 50, 50, 10, False),
         "kkt_ineq_only": (3, 4, 30, 40, 0, False), "kkt_singular": (4, 4, 20, 15, 5, True)}


def main():
    _, rb = ref_runner.load()
    os.makedirs(OUT, exist_ok=True)
    for name, (seed, B, nz, nineq, neq, sing) in CASES.items():
        pr = problem(seed, B, nz, nineq, neq, sing)
        D = torch.diag_embed(pr["d"])
        A = pr["A"] if neq else torch.Tensor().double()
        ry = pr["ry"] if neq else None
        out = {}
        try:
            full = rb.factor_solve_kkt(pr["Q"], D, pr["G"], A, pr["rx"], pr["rs"], pr["rz"], ry)
            out.update({"full_" + k: (v.numpy() if v is not None else np.zeros((B, 0))) for k, v in zip(("dx", "ds", "dz", "dy"), full)})
        except Exception as exc:      # noqa: BLE001  (singular case: LU may raise)
            print(name, "factor_solve_kkt failed:", str(exc)[:80])
        ir = rb.solve_kkt_ir(pr["Q"], D, pr["G"], A, pr["rx"], pr["rs"], pr["rz"], ry, niter=1)
        out.update({"ir_" + k: (v.numpy() if v is not None else np.zeros((B, 0))) for k, v in zip(("dx", "ds", "dz", "dy"), ir)})
        res = rb.kkt_resid_reg(pr["Q"], D, pr["G"], A, 1e-7, ir[0], ir[1], ir[2], ir[3], pr["rx"], pr["rs"], pr["rz"], ry)
        out["ir_resid_max"] = np.array(max(float(v.abs().max()) for v in res if v is not None))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: v.numpy() for k, v in pr.items()}, **out)
        print(name, "reference IR residual max %.2e" % out["ir_resid_max"], "finite", all(np.isfinite(v).all() for v in out.values()))


if __name__ == "__main__":
    main()
