// DFMA issue-rate microbenchmark: how many distinct register operands can a DFMA read at full rate?
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void __launch_bounds__(256) k(double* out, int iters, double s0, double s1) {
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 1.0 + 1e-9 * (i + threadIdx.x); c[i] = 1e-9 * i + s1; }
    double m = s0, cc = s1;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (V == 0) a[i] = fma(a[i], m, cc);            // 1 distinct varying reg + 2 shared
                if (V == 1) a[i] = fma(a[i], b[i], cc);         // 2 varying + 1 shared
                if (V == 2) a[i] = fma(b[i], c[(i + u) & 7], a[i]);   // 3 distinct regs
                if (V == 3) a[i] = fma(b[i], m, a[i]);          // acc + varying + shared (our r-inner pattern)
            }
        }
    }
    double s = 0; for (int i = 0; i < 8; i++) s += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V> void run(const char* name) {
    int grid = 148 * 8, iters = 2048; double* out; cudaMalloc(&out, sizeof(double) * grid * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0); k<V><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    double flops = 2.0 * 64 * iters * 256.0 * grid;
    printf("%s: %.3f ms  %.2f TFLOP/s\n", name, best, flops / (best * 1e-3) / 1e12);
    cudaFree(out);
}
int main() { run<0>("V0 a=fma(a,m,c)      "); run<1>("V1 a=fma(a,b_i,c)    "); run<2>("V2 a=fma(b_i,c_j,a)  "); run<3>("V3 a=fma(b_i,m,a)    "); return 0; }
