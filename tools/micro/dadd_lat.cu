// Dependent-chain latency of DADD / DFMA / (LDS + DADD) on one warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n, double inc) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = inc * (i & 7);
    __syncthreads();
    double s = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) s = s + inc;
    long long t1 = clock64();
    double f = threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < n; i++) f = fma(f, 1.0000001, inc);
    long long t2 = clock64();
    double g = 0;
#pragma unroll 8
    for (int i = 0; i < n; i++) g = g + sm[(i * 3) & 2047];
    long long t3 = clock64();
    out[threadIdx.x] = s + f + g;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 32); cudaMalloc(&cyc, 24);
    const int n = 4096;
    for (int r = 0; r < 2; r++) k<<<1, 32>>>(out, cyc, n, 1e-9);
    long long h[3]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    printf("1 warp: DADD chain %.1f cyc/op, DFMA chain %.1f cyc/op, LDS+DADD chain %.1f cyc/op\n", h[0] / (double)n, h[1] / (double)n, h[2] / (double)n);
    return 0;
}
