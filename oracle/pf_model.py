"""Numpy model of the product-form factorization and substitutions of qpth_b200/csrc/qp_pf.cuh (TEST INFRASTRUCTURE ONLY).

Restates, step for step, what the kernels do to the reduced KKT matrix S of factor_kkt / solve_kkt
(qpth/solvers/pdipm/batch.py:435-470, 349-372): the staircase storage, T_k = L_kk^-1 on the diagonal tiles, P_ik = L_ik T_k
below them, the running right-hand side, `c_k = T_k^T T_k b_k` and the backward sweep `w_k = c_k - sum P_ik^T w_i`; plus the
mirrored tile table of the trailing update. tests/test_oracle.py checks it against numpy.linalg.solve, which separates
"the algebra / indexing of the formulation is wrong" from "the CUDA code has a bug"."""
import numpy as np


def rowoff(r):
    i = r >> 3
    return (32 * i + 64) * i + (r & 7) * (8 * i + 12)


def elems(nts):
    return (32 * nts + 64) * nts


def to_staircase(M):
    n = M.shape[0]
    S = np.zeros(elems(n // 8))
    for r in range(n):
        w = 8 * (r >> 3) + 8
        S[rowoff(r):rowoff(r) + w] = np.where(np.arange(w) <= r, M[r, :w], 0.0)
    return S


def trailing_tiles(nts, k):
    """(i, j) of the tiles the update warps touch at step k, in table order (mirrored coordinates, qp_pf.cuh)."""
    out = []
    for jp in range(nts - 1 - k):
        for ip in range(jp + 1):
            out.append((nts - 1 - ip, nts - 1 - jp))
    return out[:-1], out[-1]          # the last one, (k+1, k+1), belongs to the chain warp


def factor_and_solve(Sfull, h, kb0=0):
    """Factor S (SPD, order 8 nts) in product form with the first kb0 block columns pre-factored (as pre_factor_kkt leaves
    them) and solve S w = h exactly the way the kernels do. Returns w."""
    n = Sfull.shape[0]
    nts = n // 8
    M = np.tril(Sfull).copy()
    c0 = 8 * kb0
    if kb0 > 0:                                            # pre_factor_kkt: partial Cholesky + conversion of those columns
        L11 = np.linalg.cholesky(Sfull[:c0, :c0])
        L21 = np.linalg.solve(L11, Sfull[c0:, :c0].T).T
        M[:c0, :c0] = np.tril(L11); M[c0:, :c0] = L21; M[c0:, c0:] = np.tril(Sfull[c0:, c0:] - L21 @ L21.T)
        for k in range(kb0):
            k0 = 8 * k
            T = np.linalg.inv(np.tril(M[k0:k0 + 8, k0:k0 + 8]))
            M[k0 + 8:, k0:k0 + 8] = M[k0 + 8:, k0:k0 + 8] @ T
            M[k0:k0 + 8, k0:k0 + 8] = np.tril(T)
    S = to_staircase(M)
    el = lambda r, c: S[rowoff(r) + c]

    def tile(i, j):
        return np.array([[el(8 * i + r, 8 * j + c) for c in range(8)] for r in range(8)])

    def put(i, j, V, lower_only=False):
        for r in range(8):
            for c in range(8):
                if not lower_only or c <= r:
                    S[rowoff(8 * i + r) + 8 * j + c] = V[r, c]

    b = np.array(h, dtype=np.float64)
    for k in range(kb0):                                   # pf_fwd over the pre-factored columns
        bk = b[8 * k:8 * k + 8].copy()
        for i in range(k + 1, nts):
            b[8 * i:8 * i + 8] -= tile(i, k) @ bk
    pan = np.zeros((n, 8))

    def factor_inv(k):
        D = tile(k, k)
        D = np.tril(D) + np.tril(D, -1).T
        return np.linalg.inv(np.linalg.cholesky(D))

    T = factor_inv(kb0)
    for k in range(kb0, nts):                              # pf_chol
        put(k, k, T, lower_only=True)
        for i in range(k + 1, nts):
            Li = tile(i, k) @ T.T
            pan[8 * i:8 * i + 8] = Li
            put(i, k, Li @ T)
        bk = b[8 * k:8 * k + 8].copy()
        for i in range(k + 1, nts):
            b[8 * i:8 * i + 8] -= tile(i, k) @ bk
        tiles, chain_tile = (trailing_tiles(nts, k) if k + 1 < nts else ([], None))
        for (i, j) in tiles + ([chain_tile] if chain_tile else []):
            put(i, j, tile(i, j) - pan[8 * i:8 * i + 8] @ pan[8 * j:8 * j + 8].T, lower_only=(i == j))
        if k + 1 < nts:
            T = factor_inv(k + 1)
    c = np.empty(n)
    for k in range(nts):                                   # pf_diag
        Tk = np.tril(tile(k, k))
        c[8 * k:8 * k + 8] = Tk.T @ (Tk @ b[8 * k:8 * k + 8])
    w = np.empty(n)
    acc = c.copy()
    w[n - 8:] = acc[n - 8:]
    for i in range(nts - 1, 0, -1):                        # pf_bwd
        wi = w[8 * i:8 * i + 8]
        for j in range(i):
            acc[8 * j:8 * j + 8] -= tile(i, j).T @ wi
        w[8 * i - 8:8 * i] = acc[8 * i - 8:8 * i]
    return w
