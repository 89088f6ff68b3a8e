"""Seeded synthetic QP batches (host side, numpy, fp64).

The generator follows the reference's own benchmark/test problem construction:
`prof-linear.py:64-75` (Q = L L^T + 1e-3 I with L ~ U[0,1), G ~ N(0,1),
h = G z0 + s0 with s0 ~ U[0,1) so every QP is strictly feasible) and
`test.py:50-55` for the equality part (b = A z0).  The classification-layer
pattern of `example-cls-layer.ipynb:107-111,126-130` (un-batched Q, G, h and a
batched p) is `cls_layer_problem`.

Everything is drawn from one `numpy.random.RandomState(seed)` so the same batch
can be rebuilt on any box (the GPU box has no access to the reference).
"""
import numpy as np

# The configurations BASELINE.json names (C1..C5).
CONFIGS = {
    "C1": dict(nBatch=4, nz=10, nineq=5, neq=0),
    "C2": dict(nBatch=128, nz=100, nineq=100, neq=0),
    "C3": dict(nBatch=1024, nz=50, nineq=50, neq=10),
    "C4": dict(nBatch=64, nz=200, nineq=200, neq=0),      # cls-layer pattern
    "C5": dict(nBatch=8192, nz=100, nineq=100, neq=0),    # sharded over GPUs
}


def random_qp_batch(nBatch, nz, nineq, neq=0, seed=0):
    """Dense random QP batch, all six inputs batched. Returns dict of fp64 arrays.

    Keys: Q (B,nz,nz), p (B,nz), G (B,nineq,nz), h (B,nineq), A (B,neq,nz),
    b (B,neq) (A, b have zero rows when neq == 0), dl (B,nz) ~ N(0,1) upstream
    gradient for parity runs.
    """
    rs = np.random.RandomState(seed)
    L = rs.rand(nBatch, nz, nz)
    Q = np.matmul(L, L.transpose(0, 2, 1)) + 1e-3 * np.eye(nz)
    G = rs.randn(nBatch, nineq, nz)
    z0 = rs.randn(nBatch, nz)
    s0 = rs.rand(nBatch, nineq)
    p = rs.randn(nBatch, nz)
    h = np.matmul(G, z0[:, :, None])[:, :, 0] + s0
    A = rs.randn(nBatch, neq, nz)
    b = np.matmul(A, z0[:, :, None])[:, :, 0]
    dl = rs.randn(nBatch, nz)
    return dict(Q=Q, p=p, G=G, h=h, A=A, b=b, dl=dl)


def cls_layer_problem(nBatch, nz, nineq, seed=0):
    """OptNet classification-layer pattern: shared Q, G, h; batched p.

    Q = tril(U) tril(U)^T + 1e-4 I, G ~ U(-1,1), h = G*0 + 1
    (`example-cls-layer.ipynb:107-111,126-130`).
    """
    rs = np.random.RandomState(seed)
    M = np.tril(np.ones((nz, nz)))
    L = rs.rand(nz, nz) * M
    Q = L @ L.T + 1e-4 * np.eye(nz)
    G = rs.uniform(-1.0, 1.0, size=(nineq, nz))
    h = np.ones(nineq)
    p = rs.randn(nBatch, nz)
    dl = rs.randn(nBatch, nz)
    return dict(Q=Q, p=p, G=G, h=h, A=np.zeros((0, nz)), b=np.zeros((0,)), dl=dl)


def c5_shard(rank, per_rank=1024):
    """Shard `rank` of BASELINE.json config 5 (B = 8192, nz = nineq = 100, eight shards of 1024): every shard is
    rebuilt from its own seed, so no rank (and no test) has to materialise the 1.3 GB global batch to check one."""
    return random_qp_batch(per_rank, 100, 100, 0, seed=5000 + rank)


def algorithmic_bytes_per_qp(nz, nineq, neq, itemsize=8):
    """compulsory HBM bytes per QP for fwd+bwd, all inputs batched."""
    n, m, e = nz, nineq, neq
    fwd_in = itemsize * (n * n + n + m * n + m + e * n + e)
    fwd_out = itemsize * (n + 2 * m + e)
    bwd_in = itemsize * n + fwd_out + itemsize * (n * n + m * n + e * n)
    bwd_out = fwd_in
    return dict(fwd_in=fwd_in, fwd_out=fwd_out, bwd_in=bwd_in, bwd_out=bwd_out,
                total=fwd_in + fwd_out + bwd_in + bwd_out)


def algorithmic_flops_per_qp(nz, nineq, neq, iters):
    """Cholesky-form flop count per QP for fwd+bwd."""
    n, m, e = nz, nineq, neq
    setup = n ** 3 / 3 + n * n * m + m * m * n
    if e > 0:
        setup += n * n * e + e * e * n + 2 * m * n * e + e ** 3 / 3 + e * e * m + m * m * e
    factor = m ** 3 / 3
    solve = 4 * n * n + 2 * (m + e) ** 2 + 4 * m * n + 4 * e * n
    resid = 2 * n * n + 4 * m * n + 4 * e * n
    return setup + (factor + solve) + iters * (factor + 2 * solve + resid) + (factor + solve + resid)
