// ed_mma.cu — the ED projection (edmd.cpp:89-157) as two small contractions on the FP64 tensor cores.
//
// For tetrads that share a parameter set the two passes of EDMD::calculate_ED_Forces over the eigenvector
// block E [12][3 N_a] are matrix products:
//
//   pass 1   P [12 x T]     =  E  . D          D [3 N_a x T]: displacements d = R x_c - avg_c of T tetrads   (edmd.cpp:106-110)
//   pass 2   S'[3 N_a x T]  =  E^T . P  + avg                 re-embedded coordinates                          (edmd.cpp:113-127)
//            F'[3 N_a x T]  =  E^T . (-P scaled / eval)       force in the reference frame
//
// ed_project_warp_kernel (ed.cu) evaluates them one tetrad per warp with scalar DFMAs and reads every
// coefficient of E from shared memory once per pass and tetrad: 147 KB = 1,390 shared-memory wavefronts per
// tetrad, which is what bounds it (profiles/ncu_r02_c3_step_details.txt).  Here T = 8 tetrads form the N
// dimension of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), so a coefficient fetched from shared memory is used for eight
// tetrads: 25 M shared-memory wavefronts per launch at 1e5 tetrads instead of 139 M (ncu).
//
// A CTA (8 warps, 2 CTAs per SM) stages one parameter set's E (row stride 3S + 4 doubles: both passes' fragment
// loads are conflict-free) and average structure once per run of tetrads that share the set.  Inside the run every
// WARP carries whole groups of eight tetrads, drawn from a shared counter; the warps never meet at a barrier, so
// their load, MMA and store phases drift apart and overlap.  (A CTA-cooperative form — eight warps split one group's
// K range and meet at a barrier per group — measured 0.572 ms at 1e5 tetrads against 0.49-0.50 for this one.)
//   pass 1   K = all atoms, four at a time: lane (g = lane/4, r = lane%4) loads atom 4s + r of tetrad g (one
//            16-atom half ahead of its use), forms its displacement with the reference's expressions and feeds it
//            as the B fragment of six DMMAs per s (3 planes x rows 0-7 / 8-11 of E).  Two accumulation chains (one
//            per row tile) in atom order: a fixed order, bit-stable.
//   transposition   the D fragments (row = eigenvector, columns = tetrads 2r, 2r + 1) go through a 768-byte per-warp
//            scratch and come back as pass 2's B fragments (projections of tetrad g on eigenvectors r, r + 4, r + 8);
//            the ED energy is summed in eigenvector order (edmd.cpp:153-157).
//   pass 2   8-atom M tiles; K = the 12 eigenvectors in three steps; the accumulators start from avg (S') and 0 (F');
//            the D fragment leaves lane (g, r) with atom 8 tau + g of tetrads 2r and 2r + 1, all three components,
//            which it rotates back to the laboratory frame (edmd.cpp:130-150) and stores.
// Every column of an MMA is computed independently of the others, so a tetrad's result does not depend on which
// other tetrads share its group: the results are identical for every shard layout / number of GPUs (GPU-tested).
// Same arithmetic as the other projection kernels up to the order of the twelve projection sums (~1e-15).
// Compiled with -fmad=false like ed.cu (the displacement and the back-rotation round like the CPU build).
// Measured (ncu, profiles/ncu_r02_ed_project_mma_details.txt): 0.50 ms at 1e5 tetrads; DMMA pipe 38 % + FP64 pipe 18 %
// active; DRAM traffic = the compulsory 18.6 KB per tetrad at 3.6 TB/s; a third of the warp time waits on coordinate loads.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace edmd {
namespace {

constexpr int kGroup = 8;                       // tetrads per group: the N dimension of the MMA tile
constexpr int kRows = kNevReg;                  // eigenvector rows staged (12)
constexpr int kPartRow = kRows * kGroup;        // doubles per warp of partial projections
constexpr int kSupW = kSupWordsHost;            // doubles per tetrad of superposition scratch (ed.cu)

// D (8x8) += A (8x4, row) . B (4x8, col).  Fragments (PTX ISA, mma.m8n8k4 .f64), g = lane >> 2, r = lane & 3:
//   A: one double, row g, column r;  B: one double, row r, column g;  C/D: two doubles, row g, columns 2r and 2r + 1.
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

struct MmaArgs {
    ParamSets ps;
    const int* param_id;
    const double* crd_in;  // [n][3][S] current coordinates
    double* crd;           // [n][3][S] out: the re-embedded coordinates (may alias crd_in)
    double* fed;           // [n][3][S] out
    double* e_ed;          // [n] out
    const double* sup;     // [n][kSupW] from the superposition kernel: R, shift, centroid of avg, centroid of x
    double scaled;
    const int* order;      // [count] tetrads of this launch sorted by parameter set
    int count;
    int chunk;             // tetrads per CTA (a multiple of kGroup)
    int prefetch;          // L2 prefetch of the coordinates: 0 none, 1 the whole group when it starts, 2 two 32-atom blocks ahead
};

template <int SC>   // SC: the atom stride when it is the usual 256 (shared-memory offsets fold into immediates), else 0
__global__ void __launch_bounds__(kThreads, 2) ed_project_mma_kernel(MmaArgs A) {
    extern __shared__ double msm[];
    __shared__ int sNext, sGrab;
    const unsigned full = 0xffffffffu;
    const int S = SC ? SC : A.ps.S;
    const int SP = 3 * S + 4;                   // 2 SP = 8 (mod 32) words: rows 8 banks apart
    double* sEV = msm;                          // [kRows][SP]
    double* sAV = sEV + kRows * SP;             // [3][S]
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __reduce_max_sync(full, tid >> 5);
    double* sXp = sAV + 3 * S + warp * kPartRow;   // this warp's [kRows][kGroup] scratch: projections, D -> B fragment layout
    const int g = lane >> 2, r = lane & 3;
    const int e0 = blockIdx.x * A.chunk;
    const int e1 = (e0 + A.chunk < A.count) ? e0 + A.chunk : A.count;
    const int nblk = S >> 5;                    // 32-atom blocks of a tetrad

    int base = e0;
    while (base < e1) {
        const int ps = __reduce_max_sync(full, A.param_id[A.order[base]]);
        const int na = __reduce_max_sync(full, A.ps.na[ps]), nev = __reduce_max_sync(full, A.ps.nev[ps]);
        if (tid == 0) { sNext = e1; sGrab = 0; }
        {   // stage the set's tables (rows >= NEV are zero; the arrays are zero-padded beyond na / nev)
            const double* EV = A.ps.evec + (size_t)ps * A.ps.NEV * 3 * S;
            const double* AV = A.ps.avg + (size_t)ps * 3 * S;
            const int rows = A.ps.NEV < kRows ? A.ps.NEV : kRows;
            for (int i = tid; i < kRows * 3 * S; i += kThreads) {
                const int j = i / (3 * S), c = i - j * (3 * S);
                sEV[j * SP + c] = j < rows ? EV[i] : 0.0;
            }
            for (int i = tid; i < 3 * S; i += kThreads) sAV[i] = AV[i];
        }
        __syncthreads();
        double el[3];                           // eigenvalues of the rows this lane holds as B fragments in pass 2
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int j = r + 4 * q;
            el[q] = j < nev ? A.ps.eval[(size_t)ps * A.ps.NEV + j] : 1.0;
        }

        // The run's groups are the 8-entry windows of `order` counted from `base`; the warps of the CTA draw them
        // from a shared counter and never meet at a barrier inside the run, so their load, MMA and store phases
        // drift apart and overlap.
        for (;;) {
            int k = 0;
            if (lane == 0) k = atomicAdd(&sGrab, 1);
            k = __shfl_sync(full, k, 0);
            const int ge = base + kGroup * k;
            if (ge >= e1) break;
            int my_t = 0;
            bool ok = false;
            if (lane < kGroup && ge + lane < e1) { my_t = A.order[ge + lane]; ok = A.param_id[my_t] == ps; }
            const unsigned m = __ballot_sync(full, ok) & 0xffu;
            const int ng = __ffs(~m) - 1;       // leading run of entries that share the set, 0..8
            if (ng > 0) {
                if (A.prefetch == 1) {   // the group's coordinates on their way into L2 (the loads below then see L2 latency, not HBM's)
                    const int per = 3 * S / 16;  // 128-byte lines per tetrad
                    for (int i0 = 0; i0 < kGroup * per; i0 += 32) {
                        const int i = i0 + lane;
                        const bool in = i < kGroup * per;
                        const int u = in ? i / per : 0, o = i - u * per;
                        const int tu = __shfl_sync(full, my_t, u);
                        if (in && u < ng) prefetch_l2(A.crd_in + (size_t)tu * 3 * S + o * 16);
                    }
                }
                const int tg = __shfl_sync(full, my_t, g);
                const bool colv = g < ng;

                // ---- pass 1: projections of the eight tetrads, K = all atoms, in 16-atom halves (loads one half ahead)
                double acc[2][2];               // rows 0-7 / 8-11 of P: one accumulation chain each, in atom order
                acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.0;
                {
                    double R[9], ca[3], cx[3];
#pragma unroll
                    for (int q = 0; q < 9; q++) R[q] = 0.0;
#pragma unroll
                    for (int q = 0; q < 3; q++) ca[q] = cx[q] = 0.0;
                    const double* XIN = A.crd_in;
                    if (colv) {
                        const double* SUP = A.sup + (size_t)tg * kSupW;
#pragma unroll
                        for (int q = 0; q < 9; q++) R[q] = SUP[q];
#pragma unroll
                        for (int q = 0; q < 3; q++) { ca[q] = SUP[12 + q]; cx[q] = SUP[15 + q]; }
                        XIN = A.crd_in + (size_t)tg * 3 * S;
                    }
                    auto load_half = [&](double (&x)[4][3], int hb) {
#pragma unroll
                        for (int s = 0; s < 4; s++) {
                            const int a = 16 * hb + 4 * s + r;
                            x[s][0] = x[s][1] = x[s][2] = 0.0;
                            if (colv && a < na) { x[s][0] = XIN[a]; x[s][1] = XIN[S + a]; x[s][2] = XIN[2 * S + a]; }
                        }
                    };
                    auto compute_half = [&](const double (&x)[4][3], int hb) {
#pragma unroll
                        for (int s = 0; s < 4; s++) {
                            const int a = 16 * hb + 4 * s + r;
                            double d[3] = {0.0, 0.0, 0.0};
                            if (colv && a < na) {   // edmd.cpp:89-94 (centred copies as qcprot.c:382-386)
                                const double ac0 = sAV[a] - ca[0], ac1 = sAV[S + a] - ca[1], ac2 = sAV[2 * S + a] - ca[2];
                                const double xc0 = x[s][0] - cx[0], xc1 = x[s][1] - cx[1], xc2 = x[s][2] - cx[2];
                                d[0] = R[0] * xc0 + R[1] * xc1 + R[2] * xc2 - ac0;
                                d[1] = R[3] * xc0 + R[4] * xc1 + R[5] * xc2 - ac1;
                                d[2] = R[6] * xc0 + R[7] * xc1 + R[8] * xc2 - ac2;
                            }
#pragma unroll
                            for (int c = 0; c < 3; c++) {   // edmd.cpp:106-110: rows 0-7 and rows 8-11 (+ four zero rows) of E
                                const double a_lo = sEV[g * SP + c * S + a];
                                const double a_hi = g < 4 ? sEV[(8 + g) * SP + c * S + a] : 0.0;
                                dmma(acc[0], a_lo, d[c]);
                                dmma(acc[1], a_hi, d[c]);
                            }
                        }
                    };
                    double xA[4][3], xB[4][3];
                    load_half(xA, 0);
#pragma unroll 1
                    for (int b = 0; b < nblk; b++) {
                        if (A.prefetch == 2 && b + 2 < nblk) {   // 8 tetrads x 3 planes x 2 lines of block b + 2
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int i = lane + 32 * h;
                                const int u = i < 48 ? i / 6 : 0, rem = i - 6 * u;
                                const int tu = __shfl_sync(full, my_t, u);
                                if (i < 48 && u < ng)
                                    prefetch_l2(A.crd_in + (size_t)tu * 3 * S + (rem >> 1) * S + 32 * (b + 2) + 16 * (rem & 1));
                            }
                        }
                        load_half(xB, 2 * b + 1);
                        compute_half(xA, 2 * b);
                        if (b + 1 < nblk) load_half(xA, 2 * b + 2);
                        compute_half(xB, 2 * b + 1);
                    }
                }
                // D fragments (row = eigenvector g / 8 + g, columns = tetrads 2r, 2r + 1) -> B fragments of pass 2
                __syncwarp();                   // the previous group's reads of the scratch are done
                *reinterpret_cast<double2*>(sXp + g * kGroup + 2 * r) = make_double2(acc[0][0], acc[0][1]);
                if (g < 4) *reinterpret_cast<double2*>(sXp + (8 + g) * kGroup + 2 * r) = make_double2(acc[1][0], acc[1][1]);
                __syncwarp();
                // the projections of tetrad g on eigenvectors r, r + 4, r + 8
                double P[3], W[3], term[3];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const double s = sXp[(r + 4 * q) * kGroup + g];
                    P[q] = s;
                    W[q] = -(s * A.scaled / el[q]);
                    term[q] = s * s / el[q];
                }
                {   // edmd.cpp:153-157, summed in eigenvector order
                    double en = 0.0;
#pragma unroll
                    for (int j = 0; j < kRows; j++) {
                        const double v = __shfl_sync(full, term[j >> 2], (lane & ~3) | (j & 3));
                        if (j < nev) en += v;
                    }
                    if (r == 0 && colv) A.e_ed[tg] = en * (0.5 * A.scaled);
                }

                // ---- pass 2: re-embedding and force (edmd.cpp:113-127), back to the laboratory frame (:130-150)
                const int ta = __shfl_sync(full, my_t, 2 * r), tb = __shfl_sync(full, my_t, 2 * r + 1);
                const bool va = 2 * r < ng, vb = 2 * r + 1 < ng;
                double Ra[9], sha[3], Rb[9], shb[3];
#pragma unroll
                for (int q = 0; q < 9; q++) Ra[q] = Rb[q] = 0.0;
#pragma unroll
                for (int q = 0; q < 3; q++) sha[q] = shb[q] = 0.0;
                if (va) {
                    const double* SUP = A.sup + (size_t)ta * kSupW;
#pragma unroll
                    for (int q = 0; q < 9; q++) Ra[q] = SUP[q];
#pragma unroll
                    for (int q = 0; q < 3; q++) sha[q] = SUP[9 + q];
                }
                if (vb) {
                    const double* SUP = A.sup + (size_t)tb * kSupW;
#pragma unroll
                    for (int q = 0; q < 9; q++) Rb[q] = SUP[q];
#pragma unroll
                    for (int q = 0; q < 3; q++) shb[q] = SUP[9 + q];
                }
                double* const Xa = A.crd + (size_t)ta * 3 * S;
                double* const Fa = A.fed + (size_t)ta * 3 * S;
                double* const Xb = A.crd + (size_t)tb * 3 * S;
                double* const Fb = A.fed + (size_t)tb * 3 * S;
#pragma unroll 1
                for (int b = 0; b < nblk; b++) {
#pragma unroll
                    for (int tau = 0; tau < 4; tau++) {
                        const int a = 32 * b + 8 * tau + g;
                        double sv[3][2], fv[3][2];
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            sv[c][0] = sv[c][1] = sAV[c * S + a];
                            fv[c][0] = fv[c][1] = 0.0;
                        }
#pragma unroll
                        for (int c = 0; c < 3; c++) {
#pragma unroll
                            for (int q = 0; q < 3; q++) {
                                const double ev = sEV[(4 * q + r) * SP + c * S + a];
                                dmma(sv[c], ev, P[q]);
                                dmma(fv[c], ev, W[q]);
                            }
                        }
                        if (a < na) {
                            if (va) {
                                const double s0 = sv[0][0] - sha[0], s1 = sv[1][0] - sha[1], s2 = sv[2][0] - sha[2];
                                const double f0 = fv[0][0], f1 = fv[1][0], f2 = fv[2][0];
                                Xa[a]         = Ra[0] * s0 + Ra[3] * s1 + Ra[6] * s2;
                                Xa[S + a]     = Ra[1] * s0 + Ra[4] * s1 + Ra[7] * s2;
                                Xa[2 * S + a] = Ra[2] * s0 + Ra[5] * s1 + Ra[8] * s2;
                                Fa[a]         = Ra[0] * f0 + Ra[3] * f1 + Ra[6] * f2;
                                Fa[S + a]     = Ra[1] * f0 + Ra[4] * f1 + Ra[7] * f2;
                                Fa[2 * S + a] = Ra[2] * f0 + Ra[5] * f1 + Ra[8] * f2;
                            }
                            if (vb) {
                                const double s0 = sv[0][1] - shb[0], s1 = sv[1][1] - shb[1], s2 = sv[2][1] - shb[2];
                                const double f0 = fv[0][1], f1 = fv[1][1], f2 = fv[2][1];
                                Xb[a]         = Rb[0] * s0 + Rb[3] * s1 + Rb[6] * s2;
                                Xb[S + a]     = Rb[1] * s0 + Rb[4] * s1 + Rb[7] * s2;
                                Xb[2 * S + a] = Rb[2] * s0 + Rb[5] * s1 + Rb[8] * s2;
                                Fb[a]         = Rb[0] * f0 + Rb[3] * f1 + Rb[6] * f2;
                                Fb[S + a]     = Rb[1] * f0 + Rb[4] * f1 + Rb[7] * f2;
                                Fb[2 * S + a] = Rb[2] * f0 + Rb[5] * f1 + Rb[8] * f2;
                            }
                        }
                    }
                }
            }
            if (ng < kGroup) {                  // the run (or the CTA's chunk) ends inside or before this window
                if (lane == 0 && ge + ng < e1) atomicMin(&sNext, ge + ng);
                break;
            }
        }
        __syncthreads();                        // every warp is done with this set's tables
        base = __reduce_max_sync(full, sNext);
        __syncthreads();                        // ... and has read sNext before thread 0 resets it
    }
}

size_t mma_smem_bytes(int S) { return sizeof(double) * ((size_t)kRows * (3 * S + 4) + 3 * S + kWarps * kPartRow); }

}  // namespace

bool ed_project_mma_supported(const ParamSets& ps) { return ps.NEV <= kRows && ps.S % 32 == 0 && ps.S <= kMaxStride; }

cudaError_t launch_ed_project_mma(const ParamSets& ps, const int* param_id, const double* crd_in, double* crd, double* fed,
                                  double* e_ed, const double* sup, const int* order, int sms, int count, double scaled,
                                  cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    if (!ed_project_mma_supported(ps)) return cudaErrorInvalidValue;
    cudaError_t err;
    static bool done[64] = {};                  // the opt-in to > 48 KB of dynamic shared memory is per device
    int dev = 0;
    err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    bool& configured = done[dev & 63];
    if (!configured) {
        const int most = (int)mma_smem_bytes(kMaxStride);
        err = cudaFuncSetAttribute(ed_project_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(ed_project_mma_kernel<kMaxStride>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    int chunk = (count + 2 * sms - 1) / (2 * sms);          // two resident CTAs per SM, one wave
    chunk = (chunk + kGroup - 1) / kGroup * kGroup;          // whole groups
    if (chunk < kGroup) chunk = kGroup;
    // L2 prefetch of the coordinates ahead of the loads: measured at 1e5 tetrads 0.499 (none) / 0.546 (whole group at its
    // start) / 0.480 ms (two blocks ahead), step 2.66 / 2.71 / 2.69 ms — inside the run-to-run spread; off by default
    int prefetch = 0;
    if (const char* v = getenv("EDMD_ED_MMA_PREFETCH")) prefetch = atoi(v);
    MmaArgs A{ps, param_id, crd_in, crd, fed, e_ed, sup, scaled, order, count, chunk, prefetch};
    const size_t smem = mma_smem_bytes(ps.S);
    const int grid = (count + chunk - 1) / chunk;
    if (ps.S == kMaxStride) ed_project_mma_kernel<kMaxStride><<<grid, kThreads, smem, st>>>(A);
    else                    ed_project_mma_kernel<0><<<grid, kThreads, smem, st>>>(A);
    return cudaGetLastError();
}

}  // namespace edmd
