// layout.cu — boundary transposition between the reference's per-tetrad xyz-interleaved arrays
// (Tetrad::coordinates etc., /root/reference/src/tetrad.hpp:36-60; MPI_Crds, src/mpilib.cpp:106-130)
// and the engine's SoA planes [n][3][S].  HBM-bound, one CTA per tetrad.
#include "common.cuh"
#include "kernels.h"

namespace edmd {

__global__ void __launch_bounds__(kThreads) pack_planes_kernel(const double* flat, const long long* off3,
                                                              const int* natoms, double* planes, int S) {
    const int t = blockIdx.x;
    const int na = natoms[t];
    const double* src = flat + off3[t];
    double* dst = planes + (size_t)t * 3 * S;
    for (int k = threadIdx.x; k < 3 * S; k += blockDim.x) {
        const int c = k / S, a = k - c * S;
        dst[k] = a < na ? src[3 * a + c] : 0.0;
    }
}

__global__ void __launch_bounds__(kThreads) unpack_planes_kernel(const double* planes, const long long* off3,
                                                                const int* natoms, double* flat, int S) {
    const int t = blockIdx.x;
    const int na = natoms[t];
    const double* src = planes + (size_t)t * 3 * S;
    double* dst = flat + off3[t];
    for (int k = threadIdx.x; k < 3 * na; k += blockDim.x) {
        const int a = k / 3, c = k - 3 * a;
        dst[k] = src[c * S + a];
    }
}

cudaError_t launch_pack_planes(const double* flat, const long long* off3, const int* natoms, double* planes,
                               int n, int S, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    pack_planes_kernel<<<n, kThreads, 0, st>>>(flat, off3, natoms, planes, S);
    return cudaGetLastError();
}
cudaError_t launch_unpack_planes(const double* planes, const long long* off3, const int* natoms, double* flat,
                                 int n, int S, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    unpack_planes_kernel<<<n, kThreads, 0, st>>>(planes, off3, natoms, flat, S);
    return cudaGetLastError();
}

}  // namespace edmd
