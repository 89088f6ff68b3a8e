// Fast path of the per-QP solve: everything in shared memory, compact code.
//
// Why a second set of routines next to qp_device.cuh: the first version inlined and fully unrolled every
// building block at every call site and ended up with ~20k SASS instructions (317 KB) per kernel, far
// beyond the instruction caches (ncu: 12-45 % "no_inst" stalls).  Here every building block is a
// __noinline__ function that exists once, shared-memory objects are addressed as OFFSETS (in doubles) into
// the one dynamic shared array so the compiler still emits LDS/STS, and the reduced KKT system is padded to
// a multiple of 8 (identity rows) so no block is ever partial.
//
// Reference functions implemented (qpth/solvers/pdipm/batch.py): factor_kkt :435-470 -> f_chol,
// solve_kkt :349-372 -> f_trsv_* + f_matvec_*, forward :47-207 -> k_forward_fast (qp_kernels.cu).
#pragma once
#include "qp_device.cuh"

namespace qpb {
namespace fast {

#define QPB_SMEM extern __shared__ __align__(16) double qsm[]

// Optional cycle accounting (build with -DQPB_TIMING): thread 0 of block 0 accumulates clock64() deltas per
// phase into g_tim[]; read back with qpb200_debug_timing(). Compiled out of the product build.
#ifdef QPB_TIMING
__device__ long long g_tim[64];
__device__ long long g_tlast;
#define QPB_TICK(i)                                                         \
    do {                                                                    \
        if (threadIdx.x == 0 && blockIdx.x == 0) {                          \
            const long long _t = clock64();                                 \
            g_tim[i] += _t - g_tlast;                                       \
            g_tlast = _t;                                                   \
        }                                                                   \
    } while (0)
#else
#define QPB_TICK(i) do {} while (0)
#endif

constexpr int kNT = 256;

// ---- 8x8 diagonal block helpers (always full blocks here) -------------------------------------------
// Factor the block at (k0,k0) of the matrix at offset A (ld), invert it, write L (lower), T^T (upper), dinv.
__device__ __noinline__ void f_factor_diag8(int A, int ld, int k0, int dinv) {
    QPB_SMEM;
    const int lane = threadIdx.x & 31;
    double* M = qsm + A + k0 * ld + k0;
    double Lk[36], T[36], rinv[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) Lk[QPB_LIDX(r, c)] = M[r * ld + c];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double piv = Lk[QPB_LIDX(c, c)];
        const double ri = rsqrt(piv);
        rinv[c] = ri;
        Lk[QPB_LIDX(c, c)] = piv * ri;
#pragma unroll
        for (int r = c + 1; r < 8; ++r) Lk[QPB_LIDX(r, c)] *= ri;
#pragma unroll
        for (int r = c + 1; r < 8; ++r)
#pragma unroll
            for (int cc = c + 1; cc <= r; ++cc)
                Lk[QPB_LIDX(r, cc)] = fma(-Lk[QPB_LIDX(r, c)], Lk[QPB_LIDX(cc, c)], Lk[QPB_LIDX(r, cc)]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        T[QPB_LIDX(c, c)] = rinv[c];
#pragma unroll
        for (int r = c + 1; r < 8; ++r) {
            double sacc = 0.0;
#pragma unroll
            for (int j = c; j < r; ++j) sacc = fma(Lk[QPB_LIDX(r, j)], T[QPB_LIDX(j, c)], sacc);
            T[QPB_LIDX(r, c)] = -rinv[r] * sacc;
        }
    }
    if (lane == 0) {        // one divergent region instead of 72 predicated stores (those compiled to 72 branches)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) M[r * ld + c] = Lk[QPB_LIDX(r, c)];
#pragma unroll
        for (int r = 1; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < r; ++c) M[c * ld + r] = T[QPB_LIDX(r, c)];
#pragma unroll
        for (int c = 0; c < 8; ++c) qsm[dinv + k0 + c] = rinv[c];
    }
    __syncwarp();
}

// T = L_kk^-1 from its transposed home in the upper triangle of the diagonal block.
__device__ __forceinline__ void f_load_T8(const double* M, int ld, const double* dv, double (&T)[36]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        T[QPB_LIDX(r, r)] = dv[r];
#pragma unroll
        for (int c = 0; c < r; ++c) T[QPB_LIDX(r, c)] = M[c * ld + r];
    }
}

// ---- Cholesky with register-resident trailing matrix and look-ahead (see chol_v2 in qp_device.cuh) ----
// n is a multiple of 8, (n - c0)/8 <= 13, blockDim.x == 256. aug (offset) is the right-hand side, length n.
__device__ __noinline__ void f_chol(int A, int ld, int n, int c0, int aug, int dinv, int tabo) {
    QPB_SMEM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int nts = (n - c0) >> 3;
    const int noff = (nts * (nts - 1)) / 2;
    const uint16_t* tab = reinterpret_cast<const uint16_t*>(qsm + tabo);
    double* M = qsm + A;
    double C[kCholMaxTiles][2];
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < kCholMaxTiles; ++s) {
            const bool ok = s < nts;
            const double2 v = ok ? *reinterpret_cast<const double2*>(M + (c0 + 8 * s + g) * ld + c0 + 8 * s + 2 * q)
                                 : make_double2(0.0, 0.0);
            C[s][0] = v.x; C[s][1] = v.y;
        }
        f_factor_diag8(A, ld, c0, dinv);
    } else {
#pragma unroll
        for (int s = 0; s < kCholMaxOff; ++s) {
            const int idx = s * 7 + warp - 1;
            const bool ok = idx < noff;
            const int tt = ok ? tab[idx] : 0;
            const double2 v = ok ? *reinterpret_cast<const double2*>(M + (c0 + 8 * (tt >> 8) + g) * ld + c0 + 8 * (tt & 255) + 2 * q)
                                 : make_double2(0.0, 0.0);
            C[s][0] = v.x; C[s][1] = v.y;
        }
    }
    QPB_TICK(20);   // tile load + first factor
    __syncthreads();
    QPB_TICK(21);
    for (int k = 0; k < nts; ++k) {
        const int k0 = c0 + 8 * k;
        // ---- phase A: rows below the diagonal block (and the aug row) times T^T
        {
            const int nbelow = n - k0 - 8;
            if (tid <= nbelow) {
                double T[36], a[8];
                f_load_T8(M + k0 * ld + k0, ld, qsm + dinv + k0, T);
                double* rowp = (tid < nbelow) ? (M + (k0 + 8 + tid) * ld + k0) : (qsm + aug + k0);
                const double2 v0 = *reinterpret_cast<const double2*>(rowp), v1 = *reinterpret_cast<const double2*>(rowp + 2),
                              v2 = *reinterpret_cast<const double2*>(rowp + 4), v3 = *reinterpret_cast<const double2*>(rowp + 6);
                a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y; a[4] = v2.x; a[5] = v2.y; a[6] = v3.x; a[7] = v3.y;
                double o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double acc = a[c] * T[QPB_LIDX(c, c)];
#pragma unroll
                    for (int j = 0; j < c; ++j) acc = fma(a[j], T[QPB_LIDX(c, j)], acc);
                    o[c] = acc;
                }
                *reinterpret_cast<double2*>(rowp) = make_double2(o[0], o[1]);
                *reinterpret_cast<double2*>(rowp + 2) = make_double2(o[2], o[3]);
                *reinterpret_cast<double2*>(rowp + 4) = make_double2(o[4], o[5]);
                *reinterpret_cast<double2*>(rowp + 6) = make_double2(o[6], o[7]);
            }
        }
        QPB_TICK(22);   // phase A work (thread 0)
        __syncthreads();
        QPB_TICK(23);   // barrier after A
        if (k + 1 >= nts) break;
        // ---- phase B: trailing update from registers; warp 0 looks ahead (diag tile k+1 -> factor)
        if (warp == 0) {
#pragma unroll
            for (int s = 1; s < kCholMaxTiles; ++s) {
                if (s == k + 1) {
                    const double* pr = M + (c0 + 8 * s + g) * ld + k0 + q;
                    const double a0 = pr[0], a1 = pr[4];
                    dmma884(C[s][0], C[s][1], -a0, a0);
                    dmma884(C[s][0], C[s][1], -a1, a1);
                    *reinterpret_cast<double2*>(M + (c0 + 8 * s + g) * ld + c0 + 8 * s + 2 * q) = make_double2(C[s][0], C[s][1]);
                }
            }
            __syncwarp();
            QPB_TICK(24);   // diag tile update + publish
            f_factor_diag8(A, ld, k0 + 8, dinv);
            QPB_TICK(25);   // factor + invert 8x8
#pragma unroll
            for (int s = 2; s < kCholMaxTiles; ++s) {
                if (s > k + 1 && s < nts) {
                    const double* pr = M + (c0 + 8 * s + g) * ld + k0 + q;
                    const double a0 = pr[0], a1 = pr[4];
                    dmma884(C[s][0], C[s][1], -a0, a0);
                    dmma884(C[s][0], C[s][1], -a1, a1);
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < kCholMaxOff; ++s) {
                const int idx = s * 7 + warp - 1;
                if (idx < noff) {
                    const int tt = tab[idx];
                    const int ti = tt >> 8, tj = tt & 255;
                    if (tj > k) {
                        const double* pa = M + (c0 + 8 * ti + g) * ld + k0 + q;
                        const double* pb = M + (c0 + 8 * tj + g) * ld + k0 + q;
                        const double a0 = pa[0], a1 = pa[4], b0 = pb[0], b1 = pb[4];
                        dmma884(C[s][0], C[s][1], -a0, b0);
                        dmma884(C[s][0], C[s][1], -a1, b1);
                        if (tj == k + 1)       // this tile belongs to the next panel: publish it
                            *reinterpret_cast<double2*>(M + (c0 + 8 * ti + g) * ld + c0 + 8 * tj + 2 * q) =
                                make_double2(C[s][0], C[s][1]);
                    }
                }
            }
            if (warp == 7) {                    // right-hand side: aug[j] -= P[j][:] . y
                double y[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) y[c] = qsm[aug + k0 + c];
                for (int j = k0 + 8 + lane; j < n; j += 32) {
                    const double* pr = M + j * ld + k0;
                    double acc = qsm[aug + j];
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc = fma(-pr[c], y[c], acc);
                    qsm[aug + j] = acc;
                }
            }
        }
        QPB_TICK(26);       // remaining diag tiles (warp 0)
        __syncthreads();
        QPB_TICK(27);       // barrier after B
    }
}

// ---- triangular solves with the inverted diagonal blocks (full 8-blocks, n multiple of 8) ----------------
// Forward over blocks [kbeg, kend): u[k] = solution entries; b[i >= kend] updated. b destroyed. b != u.
__device__ __noinline__ void f_trsv_fwd(int A, int ld, int n, int kbeg, int kend, int dinv, int b, int u) {
    QPB_SMEM;
    const int tid = threadIdx.x;
    const double* M = qsm + A;
    for (int k0 = kbeg; k0 < kend; k0 += 8) {
        const int nbelow = n - k0 - 8;
        if (tid < nbelow || tid == 0) {
            double T[36], y[8];
            f_load_T8(M + k0 * ld + k0, ld, qsm + dinv + k0, T);
            double r[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) r[c] = qsm[b + k0 + c];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double acc = r[c] * T[QPB_LIDX(c, c)];
#pragma unroll
                for (int j = 0; j < c; ++j) acc = fma(r[j], T[QPB_LIDX(c, j)], acc);
                y[c] = acc;
            }
            if (tid == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) qsm[u + k0 + c] = y[c];
            }
            if (tid < nbelow) {
                const double* rowp = M + (k0 + 8 + tid) * ld + k0;
                double acc = qsm[b + k0 + 8 + tid];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc = fma(-rowp[c], y[c], acc);
                qsm[b + k0 + 8 + tid] = acc;
            }
        }
        __syncthreads();
    }
}

// Backward: L^T w = u over all blocks. u destroyed. u != w.
__device__ __noinline__ void f_trsv_bwd(int A, int ld, int n, int dinv, int u, int w) {
    QPB_SMEM;
    const int tid = threadIdx.x;
    const double* M = qsm + A;
    for (int k0 = n - 8; k0 >= 0; k0 -= 8) {
        if (tid < k0 || tid == 0) {
            double T[36], y[8], r[8];
            f_load_T8(M + k0 * ld + k0, ld, qsm + dinv + k0, T);
#pragma unroll
            for (int c = 0; c < 8; ++c) r[c] = qsm[u + k0 + c];
#pragma unroll
            for (int c = 0; c < 8; ++c) {                       // y = T^T r
                double acc = r[c] * T[QPB_LIDX(c, c)];
#pragma unroll
                for (int j = c + 1; j < 8; ++j) acc = fma(r[j], T[QPB_LIDX(j, c)], acc);
                y[c] = acc;
            }
            if (tid == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) qsm[w + k0 + c] = y[c];
            }
            if (tid < k0) {
                double acc = qsm[u + tid];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc = fma(-M[(k0 + c) * ld + tid], y[c], acc);
                qsm[u + tid] = acc;
            }
        }
        __syncthreads();
    }
}

// ---- packed-L substitution (x~ = L^-1 x, x = L^-T x~): twice per kernel, off the hot loop ----------------
__device__ __noinline__ void f_whiten(int Lp, int n, int dinvL, int b, int u) {
    QPB_SMEM;
    trsv_fwd(qsm + Lp, PackedIdx{}, n, 0, n, qsm + dinvL, qsm + b, qsm + u, (int)threadIdx.x, kNT);
}
__device__ __noinline__ void f_unwhiten(int Lp, int n, int dinvL, int u, int w) {
    QPB_SMEM;
    trsv_bwd(qsm + Lp, PackedIdx{}, n, qsm + dinvL, qsm + u, qsm + w, (int)threadIdx.x, kNT);
}

// ---- mat-vecs with W (rows x cols, ld) ---------------------------------------------------------------------
// y1 = W x1, y2 = W x2 (4 lanes per row, conflict free for ld % 8 == 4)
__device__ __noinline__ void f_matvec_rows2(int W, int ld, int rows, int cols, int x1, int x2, int y1, int y2) {
    QPB_SMEM;
    matvec_rows<true>(qsm + W, ld, rows, cols, qsm + x1, qsm + x2, qsm + y1, qsm + y2, (int)threadIdx.x, kNT);
}
__device__ __noinline__ void f_matvec_rows1(int W, int ld, int rows, int cols, int x1, int y1) {
    QPB_SMEM;
    matvec_rows<false>(qsm + W, ld, rows, cols, qsm + x1, nullptr, qsm + y1, nullptr, (int)threadIdx.x, kNT);
}
// out[c] = a[c] + sgn * (W^T v)[c] (+ b[c] if b >= 0). Two row groups, partial sums in p0/p1.
__device__ __noinline__ void f_matvec_cols(int W, int ld, int rows, int cols, int v, int p0, int p1, int out,
                                           int a, double sa, int b, double sgn) {
    QPB_SMEM;
    const int tid = threadIdx.x;
    const int half = kNT / 2;
    const int gidx = tid / half, c = tid - gidx * half;
    const int chunk = (rows + 1) >> 1;
    for (int c0 = 0; c0 < cols; c0 += half) {
        const int cc = c0 + c;
        if (cc < cols) {
            const int r0 = gidx * chunk, r1 = min(rows, r0 + chunk);
            const double* Wp = qsm + W + cc;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int r = r0;
            for (; r + 3 < r1; r += 4) {
                s0 = fma(Wp[(r + 0) * ld], qsm[v + r + 0], s0);
                s1 = fma(Wp[(r + 1) * ld], qsm[v + r + 1], s1);
                s2 = fma(Wp[(r + 2) * ld], qsm[v + r + 2], s2);
                s3 = fma(Wp[(r + 3) * ld], qsm[v + r + 3], s3);
            }
            for (; r < r1; ++r) s0 = fma(Wp[r * ld], qsm[v + r], s0);
            qsm[(gidx ? p1 : p0) + cc] = (s0 + s1) + (s2 + s3);
        }
    }
    __syncthreads();
    for (int cc = tid; cc < cols; cc += kNT) {
        double r = sa * qsm[a + cc] + sgn * (qsm[p0 + cc] + qsm[p1 + cc]);
        if (b >= 0) r += qsm[b + cc];
        qsm[out + cc] = r;
    }
    __syncthreads();
}

// || L x ||^2 partial (packed lower L in shared memory)
__device__ __noinline__ double f_tri_norm2(int Lp, int n, int x) {
    QPB_SMEM;
    return tri_norm2_partial(qsm + Lp, n, qsm + x, (int)threadIdx.x, kNT);
}

__device__ __noinline__ void f_reduce_sum4(double (&v)[4], int red) {
    QPB_SMEM;
    block_reduce<4, false>(v, qsm + red, (int)threadIdx.x, kNT);
}
__device__ __noinline__ void f_reduce_sum2(double (&v)[2], int red) {
    QPB_SMEM;
    block_reduce<2, false>(v, qsm + red, (int)threadIdx.x, kNT);
}
__device__ __noinline__ void f_reduce_min2(double (&v)[2], int red) {
    QPB_SMEM;
    block_reduce<2, true>(v, qsm + red, (int)threadIdx.x, kNT);
}

}  // namespace fast
}  // namespace qpb
