// pairlist.cu — tetrad centroids and the centroid-cutoff pair list.
//
// Replaces Master::cal_Centre_of_Mass (/root/reference/src/master.cpp:157-176) and
// Master::generate_Pair_Lists (:180-210): same inclusion predicate
//     |c_i - c_j|^2 < mole_Cutoff^2  &&  |i-j| > mole_Least  &&  |i-j| < n - mole_Least,
// same orientation rule (odd separation -> (j,i), even -> (i,j)) and the same order
// (i ascending, then j ascending), so per-tetrad force sums run in the reference's order.
// The pair count is 64-bit (the reference's int overflows at n >= 46,341, master.cpp:77).
//
// The inclusion test is discontinuous, so the centroid is summed sequentially in atom order
// with the reference's expression tree (no FMA): borderline pairs fall on the same side as on
// the CPU.  Runs every ntsync steps only.
#include "common.cuh"
#include "kernels.h"

namespace edmd {

__global__ void centroid_kernel(const double* crd, const int* natoms, double* cen, int n, int S) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double* X = crd + (size_t)t * 3 * S;
    const int na = natoms[t];
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int a = 0; a < na; a++) {
        cx = __dadd_rn(cx, X[a]); cy = __dadd_rn(cy, X[S + a]); cz = __dadd_rn(cz, X[2 * S + a]);
    }
    cen[4 * (size_t)t] = cx / na; cen[4 * (size_t)t + 1] = cy / na; cen[4 * (size_t)t + 2] = cz / na;
    cen[4 * (size_t)t + 3] = 0.0;
}

__device__ __forceinline__ bool pair_included(const double* cen, int i, int j, int n, double cut2,
                                              double least) {
    const double dx = __dsub_rn(cen[4 * (size_t)i], cen[4 * (size_t)j]);
    const double dy = __dsub_rn(cen[4 * (size_t)i + 1], cen[4 * (size_t)j + 1]);
    const double dz = __dsub_rn(cen[4 * (size_t)i + 2], cen[4 * (size_t)j + 2]);
    const double r = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const int sep = j - i;
    return (r < cut2) && ((double)sep > least) && ((double)sep < ((double)n - least));
}

// One warp per row i; lanes sweep j = i+1 .. n-1 in ascending chunks of 32.
// pass 0: count; pass 1: ordered fill at row_off[i].
__global__ void pair_rows_kernel(const double* cen, int n, double cut2, double least,
                                 long long* row_count, const long long* row_off, int2* pairs, int fill) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_block) {
        long long cnt = 0;
        const long long off = fill ? row_off[i] : 0;
        for (int j0 = i + 1; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool in = (j < n) && pair_included(cen, i, j, n, cut2, least);
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (fill && in) {
                const int rank = __popc(bal & ((1u << lane) - 1u));
                const int sep = j - i;
                pairs[off + cnt + rank] = (sep & 1) ? make_int2(j, i) : make_int2(i, j);   // master.cpp:197-201
            }
            cnt += __popc(bal);
        }
        if (!fill && lane == 0) row_count[i] = cnt;
    }
}

cudaError_t launch_centroids(const double* crd, const int* natoms, double* cen, int n, int S, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    centroid_kernel<<<(n + 127) / 128, 128, 0, st>>>(crd, natoms, cen, n, S);
    return cudaGetLastError();
}

cudaError_t launch_pair_rows(const double* cen, int n, double cut2, double least, long long* row_count,
                             const long long* row_off, int2* pairs, int fill, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int wpb = 8;
    int grid = (n + wpb - 1) / wpb;
    if (grid > 148 * 16) grid = 148 * 16;
    pair_rows_kernel<<<grid, wpb * 32, 0, st>>>(cen, n, cut2, least, row_count, row_off, pairs, fill);
    return cudaGetLastError();
}

}  // namespace edmd
