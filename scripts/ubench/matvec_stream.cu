// Micro-benchmark for DESIGN.md section 10.1: can the three W passes of a Newton iteration stream W from L2
// through a small TMA ring instead of keeping all 80 KB of it in shared memory (which is what pins the solve
// kernels at one CTA per SM)?  One CTA = one "QP" with its own W (rows x cols, row-major, leading dimension ld);
// every CTA does `reps` iterations of  y = W x  (row pass, two right-hand sides as f_matvec_rows2)  followed by
// g = W^T v  (column pass, as f_matvec_cols), either from shared memory (mode 0: W staged once, the current
// design) or streamed (mode 1: CHUNK rows per stage, NSTAGE stages, 1-D bulk copies + mbarriers).
// Results are checked against a host computation; cycles per pass are reported for 1, 2 and 3 CTAs per SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o matvec_stream matvec_stream.cu && ./matvec_stream
//
// (Written on the GPU-less build container at the end of round 1; compile-checked only.)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../qpth_b200/csrc/qp_device.cuh"

using namespace qpb;

constexpr int kThreadsMV = 256;
constexpr int CHUNK = 16;       // rows per stage: 16 x 100 x 8 B = 12.8 KB
constexpr int NSTAGE = 2;

struct MVArgs {
    const double* W;            // nqp x rows x ld
    const double* x1;           // nqp x cols
    const double* x2;
    const double* v;            // nqp x rows
    double* y1;                 // nqp x rows
    double* y2;
    double* g;                  // nqp x cols
    long long* cyc;             // per CTA: cycles of the timed loop
    int rows, cols, ld, reps;
};

// ---- row pass on `nr` rows starting at row r0 of the tile at `Wt` (shared memory): 16 lanes per row
__device__ __forceinline__ void rows_tile(const double* Wt, int ld, int nr, int cols, const double* x1, const double* x2,
                                          double* y1, double* y2, int r0) {
    const int tid = threadIdx.x, rl = tid >> 4, l = tid & 15;       // 16 rows x 16 lanes
    double s1 = 0.0, s2 = 0.0;
    if (rl < nr) {
        const double* a = Wt + rl * ld;
        for (int c = l; c < cols; c += 16) {
            const double w = a[c];
            s1 = fma(w, x1[c], s1);
            s2 = fma(w, x2[c], s2);
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (rl < nr && l == 0) { y1[r0 + rl] = s1; y2[r0 + rl] = s2; }
}

// ---- column pass contribution of `nr` rows: thread c accumulates sum_r W[r][c] v[r0 + r]
__device__ __forceinline__ double cols_tile(const double* Wt, int ld, int nr, int cols, const double* v, int r0, double acc) {
    const int c = threadIdx.x;
    if (c < cols) {
#pragma unroll 4
        for (int r = 0; r < nr; ++r) acc = fma(Wt[r * ld + c], v[r0 + r], acc);
    }
    return acc;
}

template <int kMode>
__global__ void __launch_bounds__(kThreadsMV) k_mv(MVArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, qp = blockIdx.x;
    const int rows = A.rows, cols = A.cols, ld = A.ld;
    const double* Wg = A.W + (size_t)qp * rows * ld;
    // layout: [W or ring][x1][x2][v][y1][y2][g][2 mbarriers per stage]
    const int wdoubles = (kMode == 0) ? rows * ld : NSTAGE * CHUNK * ld;
    double* Ws = sm;
    double* x1 = Ws + wdoubles;
    double* x2 = x1 + cols;
    double* v = x2 + cols;
    double* y1 = v + rows;
    double* y2 = y1 + rows;
    double* g = y2 + rows;
    uint64_t* bars = reinterpret_cast<uint64_t*>(g + cols + (cols & 1));
    for (int i = tid; i < cols; i += kThreadsMV) { x1[i] = A.x1[(size_t)qp * cols + i]; x2[i] = A.x2[(size_t)qp * cols + i]; }
    for (int i = tid; i < rows; i += kThreadsMV) v[i] = A.v[(size_t)qp * rows + i];
    if (tid == 0) for (int s = 0; s < NSTAGE; ++s) mbar_init(bars + s, 1);
    __syncthreads();
    uint32_t phase[NSTAGE];
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) phase[s] = 0;
    if (kMode == 0) {
        if (tid == 0) {
            mbar_expect_tx(bars, (uint32_t)(rows * ld * 8));
            bulk_issue_thread(Ws, Wg, (uint32_t)(rows * ld * 8), bars);
        }
        mbar_wait(bars, 0);
        phase[0] = 1;
    }
    const int nchunks = (rows + CHUNK - 1) / CHUNK;
    __syncthreads();
    const long long t0 = clock64();
    for (int rep = 0; rep < A.reps; ++rep) {
        for (int pass = 0; pass < 2; ++pass) {                     // 0: rows (two vectors), 1: columns
            double acc = 0.0;
            if (kMode == 0) {
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int r0 = ch * CHUNK, nr = min(CHUNK, rows - r0);
                    if (pass == 0) rows_tile(Ws + r0 * ld, ld, nr, cols, x1, x2, y1, y2, r0);
                    else acc = cols_tile(Ws + r0 * ld, ld, nr, cols, v, r0, acc);
                }
            } else {
                // prologue: fill the ring
                if (tid == 0) {
                    for (int s = 0; s < NSTAGE && s < nchunks; ++s) {
                        const int r0 = s * CHUNK, nr = min(CHUNK, rows - r0);
                        mbar_expect_tx(bars + s, (uint32_t)(nr * ld * 8));
                        bulk_issue_thread(Ws + s * CHUNK * ld, Wg + (size_t)r0 * ld, (uint32_t)(nr * ld * 8), bars + s);
                    }
                }
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int s = ch % NSTAGE;
                    const int r0 = ch * CHUNK, nr = min(CHUNK, rows - r0);
                    mbar_wait(bars + s, phase[s]);
                    phase[s] ^= 1u;
                    const double* Wt = Ws + s * CHUNK * ld;
                    if (pass == 0) rows_tile(Wt, ld, nr, cols, x1, x2, y1, y2, r0);
                    else acc = cols_tile(Wt, ld, nr, cols, v, r0, acc);
                    __syncthreads();                               // everybody is done with stage s
                    if (tid == 0 && ch + NSTAGE < nchunks) {
                        const int r1 = (ch + NSTAGE) * CHUNK, nr1 = min(CHUNK, rows - r1);
                        fence_proxy_async();
                        mbar_expect_tx(bars + s, (uint32_t)(nr1 * ld * 8));
                        bulk_issue_thread(Ws + s * CHUNK * ld, Wg + (size_t)r1 * ld, (uint32_t)(nr1 * ld * 8), bars + s);
                    }
                }
            }
            if (pass == 1 && tid < cols) g[tid] = acc;
            __syncthreads();
        }
    }
    const long long t1 = clock64();
    for (int i = tid; i < rows; i += kThreadsMV) { A.y1[(size_t)qp * rows + i] = y1[i]; A.y2[(size_t)qp * rows + i] = y2[i]; }
    for (int i = tid; i < cols; i += kThreadsMV) A.g[(size_t)qp * cols + i] = g[i];
    if (tid == 0) A.cyc[qp] = t1 - t0;
}

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    const int rows = 100, cols = 100, ld = 100, reps = 20;
    int sms = 148;
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, 0));
    sms = prop.multiProcessorCount;
    const int maxqp = 3 * sms;
    std::vector<double> W((size_t)maxqp * rows * ld), x1((size_t)maxqp * cols), x2(x1.size()), v((size_t)maxqp * rows);
    srand(1);
    auto rnd = []() { return (double)rand() / RAND_MAX - 0.5; };
    for (auto& a : W) a = rnd();
    for (auto& a : x1) a = rnd();
    for (auto& a : x2) a = rnd();
    for (auto& a : v) a = rnd();
    MVArgs A;
    double *dW, *dx1, *dx2, *dv, *dy1, *dy2, *dg;
    long long* dcyc;
    CKC(cudaMalloc(&dW, W.size() * 8)); CKC(cudaMalloc(&dx1, x1.size() * 8)); CKC(cudaMalloc(&dx2, x2.size() * 8));
    CKC(cudaMalloc(&dv, v.size() * 8)); CKC(cudaMalloc(&dy1, v.size() * 8)); CKC(cudaMalloc(&dy2, v.size() * 8));
    CKC(cudaMalloc(&dg, x1.size() * 8)); CKC(cudaMalloc(&dcyc, maxqp * sizeof(long long)));
    CKC(cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(dx1, x1.data(), x1.size() * 8, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(dx2, x2.data(), x2.size() * 8, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(dv, v.data(), v.size() * 8, cudaMemcpyHostToDevice));
    A.W = dW; A.x1 = dx1; A.x2 = dx2; A.v = dv; A.y1 = dy1; A.y2 = dy2; A.g = dg; A.cyc = dcyc;
    A.rows = rows; A.cols = cols; A.ld = ld; A.reps = reps;
    const size_t vecs = (size_t)(3 * cols + 3 * rows + 2) * 8 + NSTAGE * 8 + 64;
    const size_t smem0 = (size_t)rows * ld * 8 + vecs, smem1 = (size_t)NSTAGE * CHUNK * ld * 8 + vecs;
    CKC(cudaFuncSetAttribute(k_mv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
    CKC(cudaFuncSetAttribute(k_mv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    printf("W %dx%d fp64 (%.1f KB per QP), %d iterations of (row pass with 2 vectors + column pass); %d SMs\n", rows, cols,
           rows * ld * 8 / 1024.0, reps, sms);
    printf("smem per CTA: resident %.1f KB, streamed %.1f KB (%d stages x %d rows)\n", smem0 / 1024.0, smem1 / 1024.0, NSTAGE, CHUNK);
    for (int mode = 0; mode < 2; ++mode) {
        for (int per_sm = 1; per_sm <= 3; ++per_sm) {
            if (mode == 0 && per_sm > 2) continue;                 // 82 KB x 3 does not fit
            const int nqp = per_sm * sms;
            cudaEvent_t e0, e1;
            CKC(cudaEventCreate(&e0)); CKC(cudaEventCreate(&e1));
            for (int it = 0; it < 3; ++it) {                       // last launch is the measured one
                CKC(cudaEventRecord(e0));
                if (mode == 0) k_mv<0><<<nqp, kThreadsMV, smem0>>>(A); else k_mv<1><<<nqp, kThreadsMV, smem1>>>(A);
                CKC(cudaEventRecord(e1));
                CKC(cudaDeviceSynchronize());
            }
            float ms = 0;
            CKC(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<long long> cyc(nqp);
            std::vector<double> y1(nqp * (size_t)rows), g(nqp * (size_t)cols);
            CKC(cudaMemcpy(cyc.data(), dcyc, nqp * sizeof(long long), cudaMemcpyDeviceToHost));
            CKC(cudaMemcpy(y1.data(), dy1, y1.size() * 8, cudaMemcpyDeviceToHost));
            CKC(cudaMemcpy(g.data(), dg, g.size() * 8, cudaMemcpyDeviceToHost));
            double err = 0.0;
            for (int q = 0; q < nqp; q += (nqp > 8 ? nqp / 8 : 1)) {
                for (int r = 0; r < rows; ++r) {
                    double s = 0; for (int c = 0; c < cols; ++c) s += W[((size_t)q * rows + r) * ld + c] * x1[(size_t)q * cols + c];
                    err = fmax(err, fabs(s - y1[(size_t)q * rows + r]));
                }
                for (int c = 0; c < cols; ++c) {
                    double s = 0; for (int r = 0; r < rows; ++r) s += W[((size_t)q * rows + r) * ld + c] * v[(size_t)q * rows + r];
                    err = fmax(err, fabs(s - g[(size_t)q * cols + c]));
                }
            }
            double mean = 0; long long mx = 0;
            for (auto c : cyc) { mean += (double)c; mx = c > mx ? c : mx; }
            mean /= nqp;
            printf("%-9s %d CTA/SM (%4d QPs): %8.0f cycles per iteration (row+col pass) mean, %8.0f max; kernel %.3f ms -> %.2f us per QP-iteration; max err %.1e\n",
                   mode == 0 ? "resident" : "streamed", per_sm, nqp, mean / reps, (double)mx / reps, ms, ms * 1e3 / reps / per_sm, err);
        }
    }
    return 0;
}
