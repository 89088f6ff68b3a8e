"""Golden-case recipes shared by `oracle/gen_golden.py` and `tests/` (TEST INFRASTRUCTURE ONLY).

Every case is rebuilt from seeds; the committed fixtures under tests/golden/
hold the reference's outputs for exactly these inputs plus an input checksum.
"""
import numpy as np

from qpth_b200.problems import random_qp_batch, cls_layer_problem


def checksum(prob):
    return float(sum(np.abs(np.asarray(prob[k], dtype=np.float64)).sum() for k in ("Q", "p", "G", "h", "A", "b")))


def proj(n):
    return np.cos(np.arange(1, n + 1, dtype=np.float64))


def shared_problem(seed=3, B=6, n=12, m=8, e=3):
    """Q, G, h, A un-batched; p, b batched -> exercises the .mean(0) rule (qp.py:159-177)."""
    pr = random_qp_batch(1, n, m, e, seed=seed)
    rs = np.random.RandomState(seed + 100)
    z0 = rs.randn(B, n)
    out = dict(Q=pr["Q"][0], G=pr["G"][0], A=pr["A"][0],
               h=pr["h"][0] + 1.0, p=rs.randn(B, n), b=z0 @ pr["A"][0].T, dl=rs.randn(B, n))
    return out


def unbatched_problem(seed=4, n=9, m=6, e=2):
    pr = random_qp_batch(1, n, m, e, seed=seed)
    return {k: v[0] for k, v in pr.items()}


def testpy_problem(nz=10, neq=2, nineq=3, Qscale=1., Gscale=1., Ascale=1.):
    """test.py:42-66 (`get_grads`): npr.seed(1), nBatch=1, loss 0.5||z - truez||^2."""
    npr = np.random.RandomState(1)
    L = npr.randn(nz, nz)
    Q = Qscale * L.dot(L.T)
    G = Gscale * npr.randn(nineq, nz)
    z0 = npr.randn(nz)
    s0 = npr.rand(nineq)
    h = G.dot(z0) + s0
    A = Ascale * npr.randn(neq, nz)
    b = A.dot(z0)
    p = npr.randn(1, nz)
    truez = npr.randn(1, nz)
    return dict(Q=Q, p=p, G=G, h=h, A=A, b=b, truez=truez)



def _testpy(tag_kw):
    def build():
        return testpy_problem(**tag_kw)
    return build


# name -> (builder, full_mats)
CASES = {
    "c1": (lambda: random_qp_batch(4, 10, 5, 0, seed=0), True),
    "eq_small": (lambda: random_qp_batch(8, 20, 15, 5, seed=1), True),
    "ineq_only_wide": (lambda: random_qp_batch(5, 7, 19, 0, seed=2), True),
    "shared": (shared_problem, True),
    "unbatched": (unbatched_problem, True),
    "c2": (lambda: random_qp_batch(128, 100, 100, 0, seed=0), False),
    "c3_b64": (lambda: random_qp_batch(64, 50, 50, 10, seed=0), False),
    "c4": (lambda: cls_layer_problem(64, 200, 200, seed=0), True),
    "c4_small": (lambda: cls_layer_problem(8, 40, 40, seed=0), True),
    "testpy_dp": (_testpy(dict(neq=2, nineq=3, Qscale=100., Gscale=100., Ascale=100.)), True),
    "testpy_dG": (_testpy(dict(neq=0, nineq=3)), True),
    "testpy_dA": (_testpy(dict(neq=3, nineq=1)), True),
}


def load_case(name, golden_dir):
    """Rebuild the inputs of golden case `name` and load the reference's outputs for it."""
    import os
    build, full_mats = CASES[name]
    prob = build()
    dlp = os.path.join(golden_dir, name + "_dl.npy")
    if os.path.exists(dlp):
        prob["dl"] = np.load(dlp)
    gold = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    cs = checksum(prob)
    assert abs(cs - float(gold["input_checksum"])) <= 1e-9 * abs(cs), \
        "golden inputs no longer reproduce from the seed (numpy RandomState changed?)"
    return prob, gold, full_mats
