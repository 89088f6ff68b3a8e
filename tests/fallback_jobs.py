"""Job list shared by tests/test_gpu_zz_fallback_families.py (parent: compares) and tests/gpu_child.py (child: solves).

The jobs exercise kernel configurations that were added to the suite after the round's last GPU session. They run in ONE
child process so that a fault or a hang in a configuration that has never run on hardware cannot take the rest of the GPU
suite (or the CUDA context of the pytest process) with it: the child is killed after a timeout, jobs that finished
before keep their results, the others fail with the reason."""

# round-1 families under QPB200_PF=0, against the goldens of the real reference
PF0_GOLDEN = ["band_smem", "band_smem_eq", "band_setup", "band_setup_eq", "c4"]

# orders above 256 (no product-form kernel, nothing fits shared memory): global-scratch family, against the oracle
BEYOND_SMEM = {
    "order_260": dict(nBatch=3, nz=120, nineq=260, neq=0, seed=41),
    "order_300": dict(nBatch=3, nz=260, nineq=300, neq=0, seed=43),
    "order_312_eq": dict(nBatch=2, nz=300, nineq=300, neq=10, seed=44),
}

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


# equality-only QPs (nineq == 0: qpth_b200/eqonly.py, an extension of the reference's surface) against the closed form
EQ_ONLY = {
    "eq_tiny": dict(B=5, nz=12, neq=5, shared=False, seed=71),          # one warp per system
    "eq_c2_sized": dict(B=4, nz=100, neq=30, shared=False, seed=72),    # product-form kernels
    "eq_shared": dict(B=6, nz=40, neq=10, shared=True, seed=73),        # un-batched Q and A: gradients are batch means
}


def eq_only_problem(B, nz, neq, shared, seed):
    import numpy as np
    rs = np.random.RandomState(seed)
    L = rs.randn(nz, nz) if shared else rs.randn(B, nz, nz)
    Q = L @ np.swapaxes(L, -1, -2) + 0.1 * np.eye(nz)
    A = rs.randn(neq, nz) if shared else rs.randn(B, neq, nz)
    return dict(Q=Q, p=rs.randn(B, nz), A=A, b=rs.randn(B, neq), dl=rs.randn(B, nz))


def jobs():
    """[(job name, kind, payload, env, mode)] in execution order."""
    out = []
    for name in PF0_GOLDEN:
        out.append(("pf0_" + name, "golden", name, {"QPB200_PF": "0"}, None))
    for name, cfg in BEYOND_SMEM.items():
        out.append(("big_" + name, "random", cfg, {}, None))
    for name, ((B, nz, nineq, neq, seed), _) in sorted(PF_OFF_GOLDEN.items()):
        cfg = dict(nBatch=B, nz=nz, nineq=nineq, neq=neq, seed=seed)
        for mode in ("latency", "throughput"):
            out.append(("pf_%s_%s" % (name, mode), "random", cfg, {}, mode))
    for name, cfg in EQ_ONLY.items():
        out.append((name, "eq_only", cfg, {}, None))
    return out
