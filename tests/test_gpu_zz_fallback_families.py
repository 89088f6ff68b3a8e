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


# Product-form configurations that no golden / sweep shape reaches (found by enumerating plan_init over a shape grid):
# (nBatch, nz, nineq, neq, seed) -> (pf_global, pf_threads, setup_pf, setup_fast, pf2_ok, pf3_ok)
PF_OFF_GOLDEN = {
    "wide_nz_eq":     ((3, 181, 49, 8, 51),  (1, 256, 1, 0, 1, 1)),   # nz > 128: W / chol(Q) from L2 at ONE QP per SM, 256 threads
    "wide_nz":        ((3, 235, 34, 0, 52),  (1, 256, 0, 0, 1, 1)),   # nz > 208: generic global-scratch setup writing the staircase
    "wide_nz_small":  ((3, 230, 20, 4, 58),  (1, 256, 0, 0, 1, 1)),
    "tall_resident":  ((3, 60, 130, 4, 54),  (0, 256, 1, 0, 0, 0)),   # order 144 > 128 with everything in shared memory
    "tall_512_eq":    ((3, 124, 190, 8, 55), (1, 512, 1, 0, 0, 0)),   # 512-thread build with equality columns
    "tall_512":       ((3, 100, 150, 0, 59), (1, 512, 1, 0, 0, 0)),
    "wide_512":       ((3, 211, 130, 0, 56), (1, 512, 0, 0, 0, 0)),   # 512-thread solve after the generic setup
    "mid_two_per_sm": ((3, 151, 100, 8, 57), (1, 256, 1, 0, 1, 0)),   # order 112: two per SM possible, three not
    "nz_above_cta":   ((3, 300, 40, 0, 60),  (1, 256, 0, 0, 1, 1)),   # nz > threads per CTA (256 and 192): strided x passes
    "nz_400_eq":      ((2, 400, 60, 8, 61),  (1, 256, 0, 0, 1, 0)),
}


@pytest.mark.parametrize("mode", ["latency", "throughput"])
@pytest.mark.parametrize("name", sorted(PF_OFF_GOLDEN))
def test_product_form_configurations_off_the_golden_shapes(name, mode, monkeypatch):
    """Each configuration against the oracle (per-QP semantics) at a well-posed shape (a quarter to a half of the
    constraints active), in latency mode and - where the shape has a several-per-SM variant - in throughput mode."""
    from qpth_b200 import _lib, qp as qpmod
    (B, nz, nineq, neq, seed), want = PF_OFF_GOLDEN[name]
    plan = _lib.plan_for(nz, nineq, neq, two=(mode == "throughput"))
    assert plan.pf == 1 and plan.tiny == 0, name
    assert (plan.pf_global, plan.pf_threads, plan.setup_pf, plan.setup_fast, plan.pf2_ok, plan.pf3_ok) == want, name
    if mode == "throughput" and not (plan.pf2_ok or plan.pf3_ok):
        pytest.skip("one QP per SM only for this shape")
    monkeypatch.setattr(qpmod, "MODE", mode)
    pr = random_qp_batch(B, nz, nineq, neq, seed=seed)
    out = _run(pr)
    ref = orc.qp_solve(pr["Q"], pr["p"], pr["G"], pr["h"], pr["A"], pr["b"], pr["dl"], per_qp=True)
    assert rel_rows(out["zhat"], ref["zhat"]).max() <= ZTOL
    for g, r in zip(out["grads"], ref["grads"]):
        if r is None:
            assert g is None
        else:
            assert rel_rows(g, r, floor=1e-4).max() <= GTOL
    assert out["iters"].max() <= 20
