// Standalone latency of the 8x8 diagonal-block factorization (single warp), variants.
#include <cstdio>
#include <cuda_runtime.h>
#define LIDX(r, c) (((r) * ((r) + 1)) / 2 + (c))
__device__ __forceinline__ double my_rsqrt(double x) {           // MUFU seed + one 3rd-order step, no special cases
    double y;
    asm("{.reg .b32 lo, hi, yh; mov.b64 {lo, hi}, %1; rsqrt.approx.ftz.f64 %0, %1;}" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, p, y);
}
template <int MODE>
__global__ void k(double* out, long long* cyc, int ld) {
    __shared__ __align__(16) double sm[8 * 12 + 64];
    int tid = threadIdx.x;
    for (int i = tid; i < 8 * 12; i += blockDim.x) { int r = i / 12, c = i % 12; sm[i] = (r == c) ? 20.0 + r : 1.0 / (1 + r + c); }
    __syncthreads();
    double Lk[36];
    long long t0 = clock64();
    double* Mb = sm;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r; c += 2) {
            const double2 v = *reinterpret_cast<const double2*>(Mb + r * ld + c);
            Lk[LIDX(r, c)] = v.x;
            if (c + 1 <= r) Lk[LIDX(r, c + 1)] = v.y;
        }
    long long t1 = clock64();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double ri;
        if (MODE == 0) ri = rsqrt(Lk[LIDX(c, c)]);
        else if (MODE == 1) ri = my_rsqrt(Lk[LIDX(c, c)]);
        else ri = 1.0 / sqrt(Lk[LIDX(c, c)]);
        Lk[LIDX(c, c)] = ri;
#pragma unroll
        for (int r = c + 1; r < 8; ++r) Lk[LIDX(r, c)] *= ri;
#pragma unroll
        for (int r = c + 1; r < 8; ++r)
#pragma unroll
            for (int cc = c + 1; cc <= r; ++cc) Lk[LIDX(r, cc)] = fma(-Lk[LIDX(r, c)], Lk[LIDX(cc, c)], Lk[LIDX(r, cc)]);
    }
    long long t2 = clock64();
    if ((tid & 31) == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c <= r; c += 2)
                *reinterpret_cast<double2*>(Mb + r * ld + c) = make_double2(Lk[LIDX(r, c)], (c + 1 <= r) ? Lk[LIDX(r, c + 1)] : 0.0);
    }
    __syncwarp();
    long long t3 = clock64();
    if (tid == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    double s = 0; for (int i = 0; i < 36; ++i) s += Lk[i];
    out[tid] = s + sm[tid & 63];
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    long long h[3];
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 32>>>(out, cyc, 12); cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost); printf("rsqrt()    : load %lld factor %lld store %lld\n", h[0], h[1], h[2]);
        k<1><<<1, 32>>>(out, cyc, 12); cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost); printf("my_rsqrt   : load %lld factor %lld store %lld\n", h[0], h[1], h[2]);
        k<2><<<1, 32>>>(out, cyc, 12); cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost); printf("1/sqrt     : load %lld factor %lld store %lld\n", h[0], h[1], h[2]);
    }
    double ho[32]; cudaMemcpy(ho, out, 256, cudaMemcpyDeviceToHost); printf("check %g\n", ho[0]);
    return 0;
}
