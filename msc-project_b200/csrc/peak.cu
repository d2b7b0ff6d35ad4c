// peak.cu — FP64 roofline denominator: DFMA issue-rate microbenchmark.
//
// MEASURED_PEAKS.json carries HBM and bf16 figures only, so the NB kernel's FP64 roofline is
// measured here: every thread runs 8 independent DFMA chains a_i = fma(a_i, b_i, c) (no memory
// traffic), 148 x 8 CTAs of 256 threads, 2 flops per DFMA.  The operand pattern matters: a DFMA
// with three distinct register sources issues at only ~76 % of this rate on B200
// (tools/micro/dfma_rf.cu: 28.3 vs 37.1 TFLOP/s), so a pattern with two register sources and
// one constant-bank source is used — the highest rate the pipe sustains.
#include "common.cuh"
#include "kernels.h"

namespace edmd {

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, long long* cycles, double c) {
    // a_i = fma(a_i, b_i, c): two register sources + one constant-bank source per DFMA
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b0 = 1.0 + 1e-9 * threadIdx.x, b1 = b0 + 1e-9, b2 = b0 + 2e-9, b3 = b0 + 3e-9;
    const double b4 = b0 + 4e-9, b5 = b0 + 5e-9, b6 = b0 + 6e-9, b7 = b0 + 7e-9;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, b0, c); a1 = fma(a1, b1, c); a2 = fma(a2, b2, c); a3 = fma(a3, b3, c);
            a4 = fma(a4, b4, c); a5 = fma(a5, b5, c); a6 = fma(a6, b6, c); a7 = fma(a7, b7, c);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}

cudaError_t measure_fp64_peak(double* tflops, double* sm_mhz_est) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8, iters = 4096;
    double* out = nullptr; long long* cyc = nullptr;
    cudaError_t err = cudaMalloc(&out, sizeof(double) * grid * 256);
    if (err != cudaSuccess) return err;
    cudaMalloc(&cyc, sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0, mhz = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        cudaMemset(cyc, 0, sizeof(long long));
        cudaEventRecord(e0);
        dfma_kernel<<<grid, 256>>>(out, iters, cyc, 1e-9);
        cudaEventRecord(e1);
        err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) break;
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * 256.0 * grid;
        const double tf = flops / (ms * 1e-3) / 1e12;
        long long c = 0; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        if (rep > 0 && tf > best) { best = tf; mhz = (double)c / (ms * 1e-3) / 1e6; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out); cudaFree(cyc);
    if (tflops) *tflops = best;
    if (sm_mhz_est) *sm_mhz_est = mhz;
    return err;
}

}  // namespace edmd
