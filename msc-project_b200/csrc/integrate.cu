// integrate.cu — Langevin/Berendsen velocity update, coordinate update and the 4-fold merge.
//
// Replaces EDMD::update_Velocities (/root/reference/src/edmd.cpp:289-317),
// EDMD::update_Coordinates (:322-327) as driven by Master::update_Velocity/update_Coordinate
// (src/master.cpp:327-343), and Master::merge_Vels_n_Crds (src/master.cpp:347-384).
//
// One CTA per tetrad, thread a <-> atom a; the velocity and coordinate updates are fused into
// a single pass over the tetrad's planes (v, F_ed, R, F_nb, m, x read once; v, x written once:
// 48,384 algorithmic bytes per tetrad).  Roofline: HBM.
#include "common.cuh"
#include "kernels.h"
#include "step_math.cuh"

namespace edmd {

__global__ void __launch_bounds__(kThreads, 4) integrate_kernel(IntArgs A) {
    __shared__ double red[1][kWarps];
    const int t = A.t0 + blockIdx.x;
    const int tid = threadIdx.x;
    const int S = A.ps.S;
    double f[3] = {0, 0, 0};
    if (A.do_vel && tid < A.ps.na[A.param_id[t]]) {
#pragma unroll
        for (int c = 0; c < 3; c++) f[c] = A.fnb[(size_t)t * 3 * S + c * S + tid];
    }
    integrate_tetrad(A, t, f, red);
}

cudaError_t launch_integrate(const ParamSets& ps, const int* param_id, double* vel, double* crd,
                             const double* fed, const double* rnd, const double* fnb, double* temp,
                             int t0, int count, int do_vel, int do_crd, const SimParams& sp,
                             cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    IntArgs A{ps, param_id, vel, crd, crd, fed, rnd, fnb, temp, t0, do_vel, do_crd,
              sp.dt, sp.gamma, sp.tautp, sp.scaled};
    integrate_kernel<<<count, kThreads, 0, st>>>(A);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Merge: every physical base pair b of the ring (0 <= b < n) belongs to the four tetrads
// b-3 .. b (indices mod n for the three wrap-around base pairs, which the .crd file repeats at
// positions n .. n+2, io.cpp:197).  The reference scatter-adds all tetrads into a flat per-bp
// array in ascending tetrad order, folds bp n+i onto bp i, multiplies by 0.25 and copies back
// (master.cpp:352-382).  Here each (bp, atom) element is owned by one thread that gathers the
// four copies in the same order, so the result is bit-identical and needs no scratch array.
struct MergeArgs {
    double* vel;
    double* crd;
    const int* bp_off;    // [n+4] prefix sum of atoms per base pair (bp_off[0] = 0)
    int n;                // tetrads
    int S;
};

__device__ __forceinline__ double gather4(double* arr, const MergeArgs& A, int b, int k, int c,
                                          bool write, double value) {
    // copies of (b,k): tetrad t covers bp t..t+3  ->  t in [b-3, b], clipped to [0, n-1]
    double sum = 0.0;
    const int n = A.n;
    if (!write) {
        if (b < 3) {
            // position n+b first (tetrads n+b-3 .. n-1), then position b (tetrads 0 .. b): the
            // fold adds flat[n+b] (left operand) and flat[b], master.cpp:366-368
            double hi = 0.0, lo = 0.0;
            for (int t = n + b - 3; t <= n - 1; t++) {
                const int a = A.bp_off[n + b] - A.bp_off[t] + k;
                hi += arr[(size_t)t * 3 * A.S + c * A.S + a];
            }
            for (int t = 0; t <= b; t++) {
                const int a = A.bp_off[b] - A.bp_off[t] + k;
                lo += arr[(size_t)t * 3 * A.S + c * A.S + a];
            }
            sum = hi + lo;
        } else {
            for (int t = b - 3; t <= b && t <= n - 1; t++) {
                const int a = A.bp_off[b] - A.bp_off[t] + k;
                sum += arr[(size_t)t * 3 * A.S + c * A.S + a];
            }
        }
        return sum;
    }
    if (b < 3) {
        for (int t = n + b - 3; t <= n - 1; t++)
            arr[(size_t)t * 3 * A.S + c * A.S + (A.bp_off[n + b] - A.bp_off[t] + k)] = value;
        for (int t = 0; t <= b; t++)
            arr[(size_t)t * 3 * A.S + c * A.S + (A.bp_off[b] - A.bp_off[t] + k)] = value;
    } else {
        for (int t = b - 3; t <= b && t <= n - 1; t++)
            arr[(size_t)t * 3 * A.S + c * A.S + (A.bp_off[b] - A.bp_off[t] + k)] = value;
    }
    return value;
}

// grid: (n base pairs), block: 3 * 64-ish threads loop over atoms x components
__global__ void __launch_bounds__(kThreads) merge_kernel(MergeArgs A) {
    const int b = blockIdx.x;
    const int nat = A.bp_off[b + 1] - A.bp_off[b];
    for (int idx = threadIdx.x; idx < 3 * nat; idx += blockDim.x) {
        const int c = idx / nat, k = idx - c * nat;
        const double v = gather4(A.vel, A, b, k, c, false, 0.0) * 0.25;
        const double x = gather4(A.crd, A, b, k, c, false, 0.0) * 0.25;
        gather4(A.vel, A, b, k, c, true, v);
        gather4(A.crd, A, b, k, c, true, x);
    }
}

cudaError_t launch_merge(double* vel, double* crd, const int* bp_off, int n, int S, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    MergeArgs A{vel, crd, bp_off, n, S};
    merge_kernel<<<n, kThreads, 0, st>>>(A);
    return cudaGetLastError();
}

}  // namespace edmd
