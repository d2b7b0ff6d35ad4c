// common.cuh — shared device helpers and the device-side data model of the EDMD engine.
//
// Data layout in HBM (all FP64, sized for 180 GB):
//   per tetrad t (dynamic, "state planes"):   crd/vel/fed/rnd/fnb  [n][3][S]
//       three SoA planes (x,y,z) of S doubles, S = atoms per tetrad rounded up to 32, so a
//       tetrad is one contiguous 3*S*8-byte block (6,144 B at S=256) that a CTA loads with
//       fully coalesced 128-bit accesses and that shared memory can hold as conflict-free
//       broadcast planes.  The reference keeps xyz-interleaved heap arrays per tetrad
//       (tetrad.hpp:36-60); the transposition happens once at the C-ABI boundary.
//   per parameter set ps (static, de-duplicated by content at upload):
//       avg [P][3][S], mass [P][3][S], chg [P][S], eval [P][NEV], evec [P][NEV][3][S]
//       param_id[n] maps tetrad -> parameter set (replicated DNA has a handful of sets, so
//       90.8 KB/tetrad of static data collapses into an L2-resident table).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edmd {

constexpr int kThreads = 256;     // one CTA = one tetrad (thread a <-> atom a)
constexpr int kWarps = kThreads / 32;
constexpr int kMaxStride = 256;   // S limit of the thread-per-atom kernels
constexpr int kNevReg = 12;       // eigenvectors held in registers per chunk

struct SimParams {                // EDMD fields, edmd.hpp:60-74 + the literals of edmd.cpp:237,242
    double dt, gamma, tautp, temperature, scaled, mole_cutoff, atom_cutoff, mole_least;
    double krep, qfac;
};

struct ParamSets {                // device pointers, de-duplicated static data
    const int* na;                // [P] atoms
    const int* nev;               // [P] eigenvectors
    const double* avg;            // [P][3][S]
    const double* mass;           // [P][3][S]
    const double* chg;            // [P][S]
    const double* eval;           // [P][NEV]
    const double* evec;           // [P][NEV][3][S]
    int S, NEV;
    const double* pre = nullptr;  // [P][4] centroid of avg + sum |avg_c|^2 (ed.cu: ps_precompute_kernel)
};

// Sum K doubles over the CTA; every thread returns with the totals in v[].
// Fixed order: xor-butterfly inside each warp, then warps 0..7 in sequence — the result is
// bit-stable from run to run and independent of scheduling.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double (*red)[kWarps]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[k][warp] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double x = red[k][0];
#pragma unroll
        for (int w = 1; w < kWarps; w++) x += red[k][w];
        v[k] = x;
    }
    __syncthreads();
}

}  // namespace edmd
