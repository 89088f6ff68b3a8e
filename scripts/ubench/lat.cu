// Microbenchmarks: latencies / throughputs of the ops the QP kernels depend on (B200, fp64).
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(int mode, double* out, long long* cyc, double seed) {
    __shared__ double sm[2048];
    int tid = threadIdx.x;
    for (int i = tid; i < 2048; i += blockDim.x) sm[i] = seed + i * 1e-3;
    __syncthreads();
    double a = seed + tid * 1e-6, b = 1.0000001, c = 1e-9, d0 = a, d1 = a + 1, d2 = a + 2, d3 = a + 3, d4 = a + 4, d5 = a + 5, d6 = a + 6, d7 = a + 7;
    int idx = tid;
    long long t0 = clock64();
    if (mode == 0) { for (int i = 0; i < N; ++i) a = fma(a, b, c); }                       // DFMA dependent chain
    else if (mode == 1) { for (int i = 0; i < N; ++i) { d0 = fma(d0, b, c); d1 = fma(d1, b, c); d2 = fma(d2, b, c); d3 = fma(d3, b, c); d4 = fma(d4, b, c); d5 = fma(d5, b, c); d6 = fma(d6, b, c); d7 = fma(d7, b, c);} a = d0+d1+d2+d3+d4+d5+d6+d7; } // 8 indep
    else if (mode == 2) { for (int i = 0; i < N; ++i) a = rsqrt(a) + 1.5; }                 // rsqrt chain (+DADD)
    else if (mode == 3) { for (int i = 0; i < N; ++i) a = 1.0 / a + 1.5; }                  // division chain
    else if (mode == 4) { for (int i = 0; i < N; ++i) { idx = (int)sm[idx & 2047] & 2047; } a = idx; }   // LDS.64 dependent (with cvt)
    else if (mode == 5) { for (int i = 0; i < N; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0; }   // SHFL.64 + DADD chain
    else if (mode == 6) { for (int i = 0; i < N; ++i) dmma(d0, d1, a, b); a = d0 + d1; }   // DMMA dependent (accumulator chain)
    else if (mode == 7) { for (int i = 0; i < N; ++i) { dmma(d0, d1, a, b); dmma(d2, d3, a, b); dmma(d4, d5, a, b); dmma(d6, d7, a, b);} a = d0+d1+d2+d3+d4+d5+d6+d7; } // 4 indep DMMA
    else if (mode == 8) { for (int i = 0; i < N; ++i) __syncthreads(); }                   // barrier
    else if (mode == 9) { for (int i = 0; i < N; ++i) a = a * b; }                         // DMUL chain
    else if (mode == 10) { for (int i = 0; i < N; ++i) a = a + c; }                        // DADD chain
    else if (mode == 11) { float f = (float)a; for (int i = 0; i < N; ++i) f = fmaf(f, 1.0001f, 1e-6f); a = f; } // FFMA chain
    else if (mode == 12) { for (int i = 0; i < N; ++i) { a = sm[(tid * 4 + i) & 2047] + a; } } // LDS + DADD dependent on acc only (throughput of LDS stream)
    else if (mode == 13) { for (int i = 0; i < N; ++i) a = sqrt(a) + 1.5; }                // sqrt chain
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = a;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const char* names[] = {"DFMA chain", "DFMA 8-indep (per 8)", "rsqrt+DADD chain", "div+DADD chain", "LDS dep (w/ cvt)", "SHFL64+DADD chain", "DMMA acc chain", "DMMA 4-indep (per 4)", "bar.sync", "DMUL chain", "DADD chain", "FFMA chain", "LDS+DADD", "sqrt+DADD chain"};
    int threads[] = {32, 64, 128, 256};
    for (int m = 0; m < 14; ++m) {
        printf("%-24s", names[m]);
        for (int ti = 0; ti < 4; ++ti) {
            k<<<1, threads[ti]>>>(m, out, cyc, 1.2345);
            cudaDeviceSynchronize();
            k<<<1, threads[ti]>>>(m, out, cyc, 1.2345);
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("  T=%3d: %7.1f", threads[ti], (double)c / N);
        }
        printf("  cyc/iter\n");
    }
    return 0;
}
