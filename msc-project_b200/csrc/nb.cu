// nb.cu — inter-tetrad non-bonded forces.
//
// Replaces EDMD::calculate_NB_Forces (/root/reference/src/edmd.cpp:232-285), the per-tetrad
// accumulation of Worker::force_Calculation (src/worker.cpp:165-179, 191-199) and the clip of
// Master::process_NB_Forces (src/master.cpp:297-322).
//
// Two kernels:
//
//  nb_scan_kernel — the FP64 hot loop.  For every listed tetrad pair it evaluates the squared
//    distance of all num_Atoms x num_Atoms atom pairs exactly as edmd.cpp:251-256 does
//    (3 DADD + DMUL + 2 DFMA per atom pair, the 9 algorithmic flops of SURVEY.md §8d) and
//    records whether ANY atom pair can contribute (r^2 below the effective cutoff).  Thread a
//    holds atom a of one tetrad in registers; the other tetrad's SoA planes sit in shared memory
//    and are read as 128-bit broadcasts (two atoms per LDS).  The cutoff test is an integer
//    min over the high words of r^2 (r^2 >= 0, so the IEEE bit pattern is monotonic) on the ALU
//    pipe, leaving the FP64 pipe only the distance arithmetic.  Output: one byte per pair.
//    Roofline: FP64 pipe.
//
//  nb_accumulate_kernel — one CTA per tetrad walks that tetrad's partners in pair-list order
//    and, for flagged pairs only, re-evaluates the pair with the reference's exact expression
//    tree (no FMA contraction, IEEE division) and the reference's summation order: inner sum
//    over the partner's atoms ascending starting from 0, then added to the tetrad total in
//    pair-list order.  Each tetrad's forces are produced by exactly one CTA in a fixed order —
//    no atomics, bit-stable, and (forces) bit-identical to the W = 1 reference.  Then the clip.
//
// With qfac = 0 (the reference's hard-coded value) an atom pair changes the result only if
// r^2 < 2 (a = max(0, 2 - r^2) > 0): in-cutoff pairs with a = 0 add +-0.0.  The effective cutoff
// is therefore min(atom_Cutoff^2, 2.0) when qfac == 0 and atom_Cutoff^2 otherwise.
#include "common.cuh"
#include "kernels.h"

namespace edmd {

__device__ __forceinline__ int hi_word(double v) { return __double2hiint(v); }

struct ScanArgs {
    const double* crd;      // [n][3][S]
    const int* natoms;      // [n] atoms per tetrad
    const int2* pairs;      // [np] (t1, t2)
    unsigned char* hit;     // [np]
    long long p0, p1;       // this launch covers pairs [p0, p1)
    int S;
    int hi_cut;             // high word of the effective squared cutoff
};

// smem: x,y,z planes of the streamed tetrad, padded to an even atom count.
__global__ void __launch_bounds__(kThreads, 4) nb_scan_kernel(ScanArgs A) {
    __shared__ __align__(16) double sx[kMaxStride];
    __shared__ __align__(16) double sy[kMaxStride];
    __shared__ __align__(16) double sz[kMaxStride];
    const int tid = threadIdx.x;
    const int S = A.S;

    for (long long p = A.p0 + blockIdx.x; p < A.p1; p += gridDim.x) {
        const int2 pr = A.pairs[p];
        // r^2 is symmetric: keep the lower-numbered tetrad in registers, stream the other.
        const int tr = pr.x < pr.y ? pr.x : pr.y;
        const int ts = pr.x < pr.y ? pr.y : pr.x;
        const int nr = A.natoms[tr], ns = A.natoms[ts];
        const double* Xr = A.crd + (size_t)tr * 3 * S;
        const double* Xs = A.crd + (size_t)ts * 3 * S;
        const int ns2 = (ns + 1) & ~1;

        double xi = 0.0, yi = 0.0, zi = 0.0;
        if (tid < nr) { xi = Xr[tid]; yi = Xr[S + tid]; zi = Xr[2 * S + tid]; }
        if (tid < ns2) {
            const bool real = tid < ns;
            sx[tid] = real ? Xs[tid] : 1e150;
            sy[tid] = real ? Xs[S + tid] : 1e150;
            sz[tid] = real ? Xs[2 * S + tid] : 1e150;
        }
        __syncthreads();

        int m0 = 0x7fffffff, m1 = 0x7fffffff;
#pragma unroll 4
        for (int j = 0; j < ns2; j += 2) {
            const double2 xs = *reinterpret_cast<const double2*>(&sx[j]);
            const double2 ys = *reinterpret_cast<const double2*>(&sy[j]);
            const double2 zs = *reinterpret_cast<const double2*>(&sz[j]);
            const double dx0 = xi - xs.x, dy0 = yi - ys.x, dz0 = zi - zs.x;
            const double dx1 = xi - xs.y, dy1 = yi - ys.y, dz1 = zi - zs.y;
            const double r0 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
            const double r1 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
            m0 = min(m0, hi_word(r0));
            m1 = min(m1, hi_word(r1));
        }
        const int pred = (tid < nr) && (min(m0, m1) <= A.hi_cut);
        const int any = __syncthreads_or(pred);   // also fences smem reuse for the next pair
        if (tid == 0) A.hit[p] = any ? 1 : 0;
    }
}

cudaError_t launch_nb_scan(const double* crd, const int* natoms, const int2* pairs, unsigned char* hit,
                           long long p0, long long p1, int S, double cut2_eff, int grid, cudaStream_t st) {
    if (p1 <= p0) return cudaSuccess;
    long long want = p1 - p0;
    if (want < grid) grid = (int)want;
    // hi-word test: r^2 < cut  ==>  hi(r^2) <= hi(cut); a superset, re-checked exactly later.
    long long bits; memcpy(&bits, &cut2_eff, 8);
    ScanArgs A{crd, natoms, pairs, hit, p0, p1, S, (int)(bits >> 32)};
    nb_scan_kernel<<<grid, kThreads, 0, st>>>(A);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
struct AccArgs {
    const double* crd;         // [n][3][S]
    const int* natoms;         // [n]
    const int* param_id;       // [n]
    const double* chg;         // [P][S] charges (abq[3i+2])
    const unsigned char* hit;  // [np]
    const long long* adj_off;  // [n+1] CSR over tetrads
    const int* adj_partner;    // partner index; bit 31 set when THIS tetrad is t2 of the pair
    const int* adj_pair;       // pair index (for the hit flag)
    double* fnb;               // [n][3][S] out, clipped
    double* e_nb;              // [n][2] out: NB energy, electrostatic energy (unclipped)
    int t0;
    int S;
    double cut2;               // atom_Cutoff^2
    double cut2_eff;           // see file header
    double krep, qfac;
    int use_hits;              // 0: treat every pair as flagged
};

__global__ void __launch_bounds__(kThreads, 2) nb_accumulate_kernel(AccArgs A) {
    __shared__ double sx[kMaxStride], sy[kMaxStride], sz[kMaxStride], sq[kMaxStride];
    __shared__ unsigned sFlag[kWarps];
    __shared__ double red[2][kWarps];

    const int t = A.t0 + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = A.S;
    const int na = A.natoms[t];
    const bool on = tid < na;
    const double* X = A.crd + (size_t)t * 3 * S;
    double xi = 0, yi = 0, zi = 0, qi = 0;
    if (on) {
        xi = X[tid]; yi = X[S + tid]; zi = X[2 * S + tid];
        qi = A.chg[(size_t)A.param_id[t] * S + tid];
    }
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;   // tetrad totals, pair-list order (worker.cpp:173-178)
    double E0 = 0.0, E1 = 0.0;             // this thread's share of the pair energies

    const long long e_begin = A.adj_off[t], e_end = A.adj_off[t + 1];
    for (long long base = e_begin; base < e_end; base += kThreads) {
        // 256 partners at a time: who is flagged?
        const long long e = base + tid;
        bool flagged = false;
        if (e < e_end) flagged = A.use_hits ? (A.hit[A.adj_pair[e]] != 0) : true;
        const unsigned bal = __ballot_sync(0xffffffffu, flagged);
        if (lane == 0) sFlag[warp] = bal;
        __syncthreads();
        unsigned flags[kWarps];
#pragma unroll
        for (int w = 0; w < kWarps; w++) flags[w] = sFlag[w];
        __syncthreads();
#pragma unroll 1
        for (int w = 0; w < kWarps; w++) {
            unsigned f = flags[w];
            while (f) {
                const int b = __ffs(f) - 1;
                f &= f - 1;
                const long long ee = base + w * 32 + b;
                const int code = A.adj_partner[ee];
                const int u = code & 0x7fffffff;
                const bool me_second = code < 0;   // this tetrad is t2 of the reference pair
                const int nu = A.natoms[u];
                const double* U = A.crd + (size_t)u * 3 * S;
                if (tid < nu) {
                    sx[tid] = U[tid]; sy[tid] = U[S + tid]; sz[tid] = U[2 * S + tid];
                    sq[tid] = A.chg[(size_t)A.param_id[u] * S + tid];
                }
                __syncthreads();
                if (on) {
                    double g0 = 0.0, g1 = 0.0, g2 = 0.0;   // per-pair partial, starts at 0 (edmd.cpp:245-246)
                    double pe0 = 0.0, pe1 = 0.0;
                    for (int j = 0; j < nu; j++) {
                        // d = x(t1) - x(t2), edmd.cpp:251-253
                        double dx, dy, dz;
                        if (me_second) { dx = __dsub_rn(sx[j], xi); dy = __dsub_rn(sy[j], yi); dz = __dsub_rn(sz[j], zi); }
                        else           { dx = __dsub_rn(xi, sx[j]); dy = __dsub_rn(yi, sy[j]); dz = __dsub_rn(zi, sz[j]); }
                        double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (r2 < 1e-9) r2 = 1e-9;
                        if (r2 < A.cut2_eff) {
                            const double a = __dsub_rn(2.0, r2) > 0.0 ? __dsub_rn(2.0, r2) : 0.0;
                            const double qq = __dmul_rn(me_second ? sq[j] : qi, me_second ? qi : sq[j]);
                            pe0 = __dadd_rn(pe0, __dmul_rn(__dmul_rn(__dmul_rn(0.25, A.krep), a), a));
                            pe1 = __dadd_rn(pe1, __dmul_rn(__dmul_rn(__dmul_rn(0.5, A.qfac), qq), r2));
                            const double pf = __dsub_rn(__dmul_rn(__dmul_rn(-2.0, A.krep), a),
                                                        __ddiv_rn(__dmul_rn(__dmul_rn(2.0, A.qfac), qq), __dmul_rn(r2, r2)));
                            if (me_second) {   // F2_j += d * pf, edmd.cpp:273-275
                                g0 = __dadd_rn(g0, __dmul_rn(dx, pf)); g1 = __dadd_rn(g1, __dmul_rn(dy, pf)); g2 = __dadd_rn(g2, __dmul_rn(dz, pf));
                            } else {           // F1_i -= d * pf, edmd.cpp:269-271
                                g0 = __dsub_rn(g0, __dmul_rn(dx, pf)); g1 = __dsub_rn(g1, __dmul_rn(dy, pf)); g2 = __dsub_rn(g2, __dmul_rn(dz, pf));
                            }
                        }
                    }
                    F0 = __dadd_rn(F0, g0); F1 = __dadd_rn(F1, g1); F2 = __dadd_rn(F2, g2);
                    E0 += pe0; E1 += pe1;
                }
                __syncthreads();
            }
        }
    }

    // clip (master.cpp:304-305) and store; energies are sums over this tetrad's atoms
    double* F = A.fnb + (size_t)t * 3 * S;
    if (on) {
        F[tid]         = F0 < -1.0 ? -1.0 : (F0 > 1.0 ? 1.0 : F0);
        F[S + tid]     = F1 < -1.0 ? -1.0 : (F1 > 1.0 ? 1.0 : F1);
        F[2 * S + tid] = F2 < -1.0 ? -1.0 : (F2 > 1.0 ? 1.0 : F2);
    } else if (tid < S) {
        F[tid] = 0.0; F[S + tid] = 0.0; F[2 * S + tid] = 0.0;
    }
    double ev[2] = {E0, E1};
    block_sum<2>(ev, red);
    if (tid == 0) { A.e_nb[2 * (size_t)t] = ev[0]; A.e_nb[2 * (size_t)t + 1] = ev[1]; }
}

cudaError_t launch_nb_accumulate(const double* crd, const int* natoms, const int* param_id, const double* chg,
                                 const unsigned char* hit, const long long* adj_off, const int* adj_partner,
                                 const int* adj_pair, double* fnb, double* e_nb, int t0, int count, int S,
                                 const SimParams& sp, int use_hits, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    const double cut2 = sp.atom_cutoff * sp.atom_cutoff;
    AccArgs A{crd, natoms, param_id, chg, hit, adj_off, adj_partner, adj_pair, fnb, e_nb, t0, S,
              cut2, nb_effective_cut2(sp), sp.krep, sp.qfac, use_hits};
    nb_accumulate_kernel<<<count, kThreads, 0, st>>>(A);
    return cudaGetLastError();
}

}  // namespace edmd
