"""GPU parity of the kernel families the product-form kernels shadow at the golden shapes, and of orders beyond them.

The default plan sends every golden case with an order of 8..256 that fits shared memory through the product-form
kernels (qp_pf.cuh). What is left for the round-1 kernels in the shipped configuration is (a) tiny problems (one warp
per QP; covered by the c1 / testpy cases of test_gpu_parity.py) and (b) problems whose factor does not fit shared
memory at all - order neq_pad + nineq > 256, or a staircase + vectors above 227 KB - which run the GLOBAL-SCRATCH
kernels (plan.smem_resident == 0, plan.pf == 0). This file pins (b) against the oracle, and re-runs the band / C4
goldens with QPB200_PF=0 so that the generic shared-memory and global-scratch solve kernels keep a parity check at the
shapes of the real reference's outputs too. (Named to sort after the other GPU files. Written after the round's GPU budget was spent: the test logic was dry-run here
with oracle/kernel_model.py standing in for the CUDA path; its first run on hardware is the round-end run.)
"""
import numpy as np
import pytest

from oracle import pdipm_oracle as orc
from oracle.cases import load_case
from qpth_b200.problems import random_qp_batch
from tests.parity import check_against_golden, rel_rows, ZTOL, GTOL
from tests.test_gpu_parity import _run, _report, EXPECTED_PATH

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["band_smem", "band_smem_eq", "band_setup", "band_setup_eq", "c4"])
def test_round1_kernel_families_match_golden(name, golden_dir, monkeypatch):
    """QPB200_PF=0: generic shared-memory kernels (nineq > 104), fast solve + generic setup (nz > 104), and the
    global-scratch kernels (C4, 200 x 200) against the real reference's outputs."""
    from qpth_b200 import _lib
    monkeypatch.setenv("QPB200_PF", "0")
    prob, gold, full = load_case(name, golden_dir)
    Qs, Gs, As = np.asarray(prob["Q"]), np.asarray(prob["G"]), np.asarray(prob["A"])
    plan = _lib.plan_for(Qs.shape[-1], Gs.shape[-2], As.shape[-2] if As.size else 0)
    assert plan.pf == 0 and (plan.fast, plan.setup_fast, plan.smem_resident) == EXPECTED_PATH[name], name
    if name == "c4":
        assert plan.solve_scratch_elems > 0 and plan.setup_scratch_elems > 0
    out = _run(prob)
    errs = check_against_golden(out, gold, full, what=name + "[pf0]", prob=prob)
    _report(name + "[pf0]", errs)


@pytest.mark.parametrize("cfg", [dict(nBatch=3, nz=120, nineq=260, neq=0, seed=41),
                                 dict(nBatch=3, nz=260, nineq=300, neq=0, seed=43),
                                 dict(nBatch=2, nz=300, nineq=300, neq=10, seed=44)])
def test_orders_beyond_shared_memory_vs_oracle(cfg):
    """Orders above 256 (no product-form kernel, nothing fits shared memory): the shipped plan is the global-scratch
    family. z* and every gradient against the oracle (per-QP semantics); the shapes are well posed (110-150 active
    constraints, so no gradient is vertex noise)."""
    from qpth_b200 import _lib
    plan = _lib.plan_for(cfg["nz"], cfg["nineq"], cfg["neq"])
    assert (plan.tiny, plan.pf, plan.smem_resident) == (0, 0, 0) and plan.solve_scratch_elems > 0
    pr = random_qp_batch(**cfg)
    out = _run(pr)
    ref = orc.qp_solve(pr["Q"], pr["p"], pr["G"], pr["h"], pr["A"], pr["b"], pr["dl"], per_qp=True)
    assert rel_rows(out["zhat"], ref["zhat"]).max() <= ZTOL
    for g, r in zip(out["grads"], ref["grads"]):
        if r is None:
            assert g is None
        else:
            assert rel_rows(g, r, floor=1e-4).max() <= GTOL
    assert out["iters"].max() <= 20
