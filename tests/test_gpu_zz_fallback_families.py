"""GPU parity of the kernel families the product-form kernels shadow at the golden shapes, of orders beyond them, and of
product-form configurations no golden shape reaches.

The default plan sends every golden case with an order of 8..256 that fits shared memory through the product-form
kernels (qp_pf.cuh). What is left for the round-1 kernels in the shipped configuration is (a) tiny problems (one warp
per QP; covered by the c1 / testpy cases of test_gpu_parity.py) and (b) problems whose factor does not fit shared
memory at all - order neq_pad + nineq > 256, or a staircase + vectors above 227 KB - which run the GLOBAL-SCRATCH
kernels (plan.smem_resident == 0, plan.pf == 0). This file pins (b) against the oracle, re-runs the band / C4 goldens
with QPB200_PF=0 so that the generic shared-memory and global-scratch solve kernels keep a parity check at the shapes of
the real reference's outputs too, and covers the product-form configurations found by enumerating plan_init over a
shape grid (wide nz, orders above 128, the 512-thread build with equalities, nz above the CTA size).

Written after the round's GPU budget was spent: the comparison logic was dry-run here with oracle/kernel_model.py
standing in for the CUDA path, and the first run on hardware is the round-end one. For that reason the solves run in
ONE child process (tests/gpu_child.py, jobs in tests/fallback_jobs.py) with a timeout: a fault or a hang in a
configuration that has never run cannot take the rest of the GPU suite with it. (Named to sort after the other files.)
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pdipm_oracle as orc
from oracle.cases import load_case
from qpth_b200.problems import random_qp_batch
from tests.fallback_jobs import BEYOND_SMEM, EQ_ONLY, PF0_GOLDEN, PF_OFF_GOLDEN, eq_only_problem
from tests.parity import check_against_golden, rel_rows, ZTOL, GTOL

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD_TIMEOUT_S = 240
# (fast, setup_fast, smem_resident) under QPB200_PF=0 (the table of test_gpu_parity.py)
PF0_PATH = {"band_smem": (0, 0, 1), "band_smem_eq": (0, 0, 1), "band_setup": (1, 0, 1), "band_setup_eq": (1, 0, 1),
            "c4": (0, 0, 0)}


@pytest.fixture(scope="module")
def child_results(tmp_path_factory):
    out_dir = str(tmp_path_factory.mktemp("fallback_families"))
    note = ""
    try:
        r = subprocess.run([sys.executable, "-m", "tests.gpu_child", out_dir], cwd=ROOT, timeout=CHILD_TIMEOUT_S,
                           capture_output=True, text=True)
        if r.returncode != 0:
            note = "child exited with %d: %s" % (r.returncode, (r.stderr or "")[-2000:])
    except subprocess.TimeoutExpired:
        note = "child killed after %d s (a job hung)" % CHILD_TIMEOUT_S
    return out_dir, note


def _load(child_results, job):
    out_dir, note = child_results
    path = os.path.join(out_dir, job + ".npz")
    if not os.path.exists(path):
        err = os.path.join(out_dir, job + ".err")
        why = open(err).read()[-3000:] if os.path.exists(err) else ("no result: " + (note or "job never ran"))
        pytest.fail("%s: %s" % (job, why), pytrace=False)
    d = np.load(path)
    out = {k: d[k] for k in ("zhat", "lam", "slacks", "iters") if k in d}
    out["nus"] = d["nus"] if "nus" in d else None
    out["grads"] = tuple(d["grad%d" % i] if ("grad%d" % i) in d else None for i in range(6))
    return out


def _plan_with_env(env, *shape, **kw):
    from qpth_b200 import _lib
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return _lib.plan_for(*shape, **kw)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _check_vs_oracle(out, pr):
    ref = orc.qp_solve(pr["Q"], pr["p"], pr["G"], pr["h"], pr["A"], pr["b"], pr["dl"], per_qp=True)
    assert rel_rows(out["zhat"], ref["zhat"]).max() <= ZTOL
    for g, r in zip(out["grads"], ref["grads"]):
        if r is None:
            assert g is None
        else:
            assert rel_rows(g, r, floor=1e-4).max() <= GTOL
    assert out["iters"].max() <= 20


@pytest.mark.parametrize("name", PF0_GOLDEN)
def test_round1_kernel_families_match_golden(name, golden_dir, child_results):
    """QPB200_PF=0: generic shared-memory kernels (nineq > 104), fast solve + generic setup (nz > 104), and the
    global-scratch kernels (C4, 200 x 200) against the real reference's outputs."""
    prob, gold, full = load_case(name, golden_dir)
    Qs, Gs, As = np.asarray(prob["Q"]), np.asarray(prob["G"]), np.asarray(prob["A"])
    plan = _plan_with_env({"QPB200_PF": "0"}, Qs.shape[-1], Gs.shape[-2], As.shape[-2] if As.size else 0)
    assert plan.pf == 0 and (plan.fast, plan.setup_fast, plan.smem_resident) == PF0_PATH[name], name
    if name == "c4":
        assert plan.solve_scratch_elems > 0 and plan.setup_scratch_elems > 0
    out = _load(child_results, "pf0_" + name)
    check_against_golden(out, gold, full, what=name + "[pf0]", prob=prob)


@pytest.mark.parametrize("name", sorted(BEYOND_SMEM))
def test_orders_beyond_shared_memory_vs_oracle(name, child_results):
    """Orders above 256 (no product-form kernel, nothing fits shared memory): the shipped plan is the global-scratch
    family. z* and every gradient against the oracle (per-QP semantics); the shapes are well posed (110-150 active
    constraints, so no gradient is vertex noise)."""
    from qpth_b200 import _lib
    cfg = BEYOND_SMEM[name]
    plan = _lib.plan_for(cfg["nz"], cfg["nineq"], cfg["neq"])
    assert (plan.tiny, plan.pf, plan.smem_resident) == (0, 0, 0) and plan.solve_scratch_elems > 0
    _check_vs_oracle(_load(child_results, "big_" + name), random_qp_batch(**cfg))


@pytest.mark.parametrize("mode", ["latency", "throughput"])
@pytest.mark.parametrize("name", sorted(PF_OFF_GOLDEN))
def test_product_form_configurations_off_the_golden_shapes(name, mode, child_results):
    """Each configuration against the oracle (per-QP semantics) at a well-posed shape (a quarter to a half of the
    constraints active), in latency mode and - where the shape has a several-per-SM variant - in throughput mode."""
    from qpth_b200 import _lib
    (B, nz, nineq, neq, seed), want = PF_OFF_GOLDEN[name]
    plan = _lib.plan_for(nz, nineq, neq, two=(mode == "throughput"))
    assert plan.pf == 1 and plan.tiny == 0, name
    assert (plan.pf_global, plan.pf_threads, plan.setup_pf, plan.setup_fast, plan.pf2_ok, plan.pf3_ok) == want, name
    if mode == "throughput" and not (plan.pf2_ok or plan.pf3_ok):
        pytest.skip("one QP per SM only for this shape")
    _check_vs_oracle(_load(child_results, "pf_%s_%s" % (name, mode)), random_qp_batch(B, nz, nineq, neq, seed=seed))


@pytest.mark.parametrize("name", sorted(EQ_ONLY))
def test_equality_only_qp_vs_closed_form(name, child_results):
    """nineq == 0 (qpth_b200/eqonly.py: two stand-alone KKT solves on the kernels): z* and the gradients against torch
    autograd through the dense KKT solution on the CPU, with the reference's conventions (symmetrised dQ, batch mean for
    un-batched inputs)."""
    import torch
    cfg = EQ_ONLY[name]
    pr = eq_only_problem(**cfg)
    B, nz, neq, shared = cfg["B"], cfg["nz"], cfg["neq"], cfg["shared"]
    Q, p, A, b = (torch.tensor(pr[k], requires_grad=True) for k in ("Q", "p", "A", "b"))
    Qb = Q.expand(B, nz, nz) if shared else Q
    Ab = A.expand(B, neq, nz) if shared else A
    K = torch.cat([torch.cat([Qb, Ab.transpose(1, 2)], 2), torch.cat([Ab, torch.zeros(B, neq, neq, dtype=Q.dtype)], 2)], 1)
    z = torch.linalg.solve(K, torch.cat([-p, b], 1).unsqueeze(-1)).squeeze(-1)[:, :nz]
    z.backward(torch.tensor(pr["dl"]))
    out = _load(child_results, name)
    assert rel_rows(out["zhat"], z.detach().numpy()).max() <= ZTOL
    ref = (0.5 * (Q.grad + Q.grad.transpose(-1, -2)) / (B if shared else 1), p.grad, A.grad / (B if shared else 1), b.grad)
    for g, r in zip(out["grads"][:4], ref):
        assert g.shape == tuple(r.shape)
        assert rel_rows(g, r.numpy(), floor=1e-4).max() <= GTOL
