// latency microbenchmarks (single warp): dependent chains, cycles per op
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = a + threadIdx.x; sm[32 + threadIdx.x] = b;
    __syncwarp();
    double x = a + threadIdx.x * 1e-9;
    long long t0, t1;
    const int R = 2048;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) x = fma(x, b, a);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) x = x + b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 4 independent DFMA chains (throughput-ish for one warp)
    double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) { y0 = fma(y0, b, a); y1 = fma(y1, b, a); y2 = fma(y2, b, a); y3 = fma(y3, b, a); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    x = y0 + y1 + y2 + y3;
    // shuffle chain (64-bit)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // LDS chain (pointer chasing through values)
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) { double v = sm[idx & 63]; idx = (int)v & 31; x += v * 0.0; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 256; i++) x = 1.0 / (x + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // rsqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 256; i++) x = rsqrt(x + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // sqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < 256; i++) x = sqrt(x + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // IMAD/int chain
    int q = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < R; i++) q = q * 3 + 1;
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    out[threadIdx.x] = x + q + idx;
}
int main()
{
    double* out; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 128);
    k<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); k<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    long long h[9]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char* n[9] = {"DFMA dep", "DADD dep", "DFMA x4 indep (per 4)", "SHFL.64 dep", "LDS.64 dep", "1.0/x dep", "rsqrt dep", "sqrt dep", "IMAD dep"};
    const int R[9] = {2048, 2048, 2048, 2048, 2048, 256, 256, 256, 2048};
    for (int i = 0; i < 9; i++) printf("%-26s %7.1f cycles/op\n", n[i], (double)h[i] / R[i]);
    return 0;
}
