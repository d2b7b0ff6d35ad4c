// rng.cu — Langevin random terms with the reference's exact pseudo-random stream.
//
// Replaces EDMD::calculate_Random_Terms (/root/reference/src/edmd.cpp:170-228): per tetrad,
// srand((unsigned)(RNG_Seed++ + rank^3 + time)) then 3*num_Atoms normal deviates from the
// ACM-712 Kinderman-Monahan ratio-of-uniforms loop fed by libc rand(), each scaled by
// sqrt(2*gamma*kT*m/dt).
//
// libc rand() is third-party (glibc TYPE_3 additive feedback generator, stdlib/random_r.c,
// call sites edmd.cpp:179,206,207).  With Y the state rotated so that Y[l] = r[(l+3) % 31],
// the generator is Y[m] = Y[m-31] + Y[m-3] (mod 2^32), output (Y[m] >> 1), and srand() discards
// the first 310 outputs.  One warp owns one tetrad's stream: lane l holds element l of the
// current 31-word block and the next block is N[l] = P[l] + N[l-3], i.e. a stride-3 inclusive
// scan over lanes (four shuffles) plus P[28 + l%3] — no serial loop over the 2,070 uniforms.
// Every Kinderman-Monahan trial consumes exactly two uniforms, accepted or not, so trial k of
// the stream uses outputs 2k, 2k+1: 31 trials (two blocks) are tested in parallel and the
// accepted ones are numbered with a ballot prefix count.  The result is the reference's
// sequence, element for element.
#include "common.cuh"
#include "kernels.h"

namespace edmd {

__device__ __forceinline__ unsigned lfg_next_block(unsigned P, int lane) {
    // stride-3 inclusive scan of P over lanes 0..30
    unsigned s = P;
#pragma unroll
    for (int d = 3; d < 31; d <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += up;
    }
    const unsigned tail = __shfl_sync(0xffffffffu, P, 28 + lane % 3);
    return s + tail;   // lane 31 carries garbage and is never read
}

// x / RAND_MAX (edmd.cpp:206-207) for an integer 0 <= x < 2^31 without the division sequence:
// q0 = x * RN(1/M), one exact-residual correction.  The result equals the correctly rounded
// quotient x / 2147483647.0 for EVERY such x (exhaustive check: tests/cpp/div_rand_max.c, run by
// tests/test_oracle.py) — 3 FP64 instructions instead of ~25.
__device__ __forceinline__ double div_rand_max(double x) {
    const double M = 2147483647.0, c = 1.0 / 2147483647.0;
    const double q0 = __dmul_rn(x, c);
    const double r = __fma_rn(-q0, M, x);
    return __fma_rn(r, c, q0);
}

// 16807^i mod (2^31 - 1), i = 0..30.  srand() fills r[i] = 16807 * r[i-1] mod (2^31-1) by Schrage's method
// (stdlib/random_r.c: hi = w / 127773, lo = w % 127773, w = 16807*lo - 2836*hi, +2^31-1 if negative), which is the
// exact modular product — also for a negative (int32) seed with C's truncating division — so r[i] is the closed
// form kSeedPow[i] * (int32)seed mod (2^31-1) and every lane gets its word without the 30-step serial loop.
// (Equality with the iteration checked for edge seeds 0, 1, 2^31-1, 2^31, 2^32-1 and 20,000 random ones:
// tests/test_oracle.py::test_glibc_seed_closed_form.)
__constant__ unsigned kSeedPow[31] = {
    1u, 16807u, 282475249u, 1622650073u, 984943658u, 1144108930u, 470211272u, 101027544u, 1457850878u, 1458777923u,
    2007237709u, 823564440u, 1115438165u, 1784484492u, 74243042u, 114807987u, 1137522503u, 1441282327u, 16531729u,
    823378840u, 143542612u, 896544303u, 1474833169u, 1264817709u, 1998097157u, 1817129560u, 1131570933u, 197493099u,
    1404280278u, 893351816u, 1505795335u};

struct RngArgs {
    ParamSets ps;
    const double* noise; // [P][3][S] sqrt(2*gamma*kT*m/dt), launch_noise_table
    const int* param_id;
    double* rnd;         // [n][3][S]
    int t0, count;
    double gamma, scaled, dt;
    long time0;
    int rank;
    long calls_before;   // calculate_Random_Terms calls made before tetrad t0 of this launch
    const long* calls_dev;   // when non-null: *calls_dev + t0 replaces calls_before (CUDA-graph replay)
};

__global__ void __launch_bounds__(128) random_terms_kernel(RngArgs A) {
    const int lane = threadIdx.x & 31;
    // warps walk the tetrads with a grid stride: a full grid gives one tetrad per warp; the step's side branch
    // launches a small resident grid instead, which leaves most of every SM to the kernels it runs beside
    // (warp index and normal count pass through redux.sync, whose result lives in a uniform register: the loops
    // below then branch on provably warp-uniform values and the shuffles in them need no reconvergence code)
    const int warps = gridDim.x * (blockDim.x >> 5);
    const int w0 = __reduce_max_sync(0xffffffffu, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    for (int w = w0; w < A.count; w += warps) {
    const int t = A.t0 + w;
    const int ps = A.param_id[t];
    const int S = A.ps.S;
    const int n3 = __reduce_max_sync(0xffffffffu, 3 * A.ps.na[ps]);
    const double* NZ = A.noise + (size_t)ps * 3 * S;
    double* out = A.rnd + (size_t)t * 3 * S;

    // seed: RNG_Seed starts at 13579, ++ per call, reset to 13579 once it exceeds 50,000,000
    // (edmd.cpp:173,179-180): the value used by call number c (0-based) is 13579 + c % 49986422.
    const long c = (A.calls_dev ? *A.calls_dev + A.t0 : A.calls_before) + w;
    const unsigned rs = 13579u + (unsigned)(c % 49986422L);
    unsigned seed = (unsigned)((long)(rs + (unsigned)(A.rank * A.rank * A.rank)) + A.time0);
    if (seed == 0) seed = 1;

    // srand(): r[0] = seed, r[i] = 16807 * r[i-1] mod 2147483647 (Schrage), i = 1..30 — closed form, see kSeedPow
    const int want = (lane + 3) % 31;   // lane l keeps r[(l+3) % 31]
    unsigned P = seed;
    if (want != 0) {
        long long rem = ((long long)kSeedPow[want] * (long long)(int)seed) % 2147483647LL;   // sign of the dividend
        if (rem < 0) rem += 2147483647LL;
        P = (unsigned)rem;
    }
    for (int b = 0; b < 10; b++) P = lfg_next_block(P, lane);   // 310 discarded outputs

    int k = 0;   // normals produced so far (warp-uniform)
    while (k < n3) {
        const unsigned Ab = lfg_next_block(P, lane);
        const unsigned Bb = lfg_next_block(Ab, lane);
        P = Bb;
        // trial `lane` (0..30) uses stream elements 2*lane, 2*lane+1 of the 62 just produced
        const int e0 = 2 * lane, e1 = 2 * lane + 1;
        const unsigned a0 = __shfl_sync(0xffffffffu, Ab, e0 < 31 ? e0 : 0);
        const unsigned b0 = __shfl_sync(0xffffffffu, Bb, e0 >= 31 ? (e0 - 31) & 31 : 0);
        const unsigned a1 = __shfl_sync(0xffffffffu, Ab, e1 < 31 ? e1 : 0);
        const unsigned b1 = __shfl_sync(0xffffffffu, Bb, e1 >= 31 ? (e1 - 31) & 31 : 0);
        const unsigned o0 = ((e0 < 31 ? a0 : b0) >> 1) & 0x7fffffffu;
        const unsigned o1 = ((e1 < 31 ? a1 : b1) >> 1) & 0x7fffffffu;

        bool accept = false;
        double z = 0.0;
        if (lane < 31) {   // edmd.cpp:206-216
            const double u = div_rand_max((double)(int)o0);
            double v = div_rand_max((double)(int)o1);
            v = __dmul_rn(1.7156, __dsub_rn(v, 0.5));
            const double x = __dsub_rn(u, 0.449871);
            const double y = __dadd_rn(fabs(v), 0.386595);
            const double q = __dadd_rn(__dmul_rn(x, x),
                                       __dmul_rn(y, __dsub_rn(__dmul_rn(0.19600, y), __dmul_rn(0.25472, x))));
            if (q < 0.27597) accept = true;
            else if (!(q > 0.27846))
                accept = __dmul_rn(v, v) < __dmul_rn(__dmul_rn(-4.0, log(u)), __dmul_rn(u, u));
            z = v / u;
        }
        const unsigned acc = __ballot_sync(0xffffffffu, accept);
        if (accept) {
            const int idx = k + __popc(acc & ((1u << lane) - 1u));
            if (idx < n3) {
                const int a = idx / 3, cc = idx - 3 * a;
                out[cc * S + a] = __dmul_rn(z, NZ[cc * S + a]);                   // edmd.cpp:223
            }
        }
        k += __popc(acc);
    }
    }   // tetrads of this warp
}

// noise[p][c][a] = sqrt(2*gamma*kT*m/dt), edmd.cpp:184 — one table per parameter set instead of one
// square root per deviate (IEEE operations in the reference's order, so the same bits).
__global__ void noise_table_kernel(const double* mass, double* noise, long total, double gamma, double scaled, double dt) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) noise[i] = sqrt(__ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, gamma), scaled), mass[i]), dt));
}
cudaError_t launch_noise_table(const ParamSets& ps, int P, double* noise, const SimParams& sp, cudaStream_t st) {
    const long total = (long)P * 3 * ps.S;
    if (total <= 0) return cudaSuccess;
    noise_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ps.mass, noise, total, sp.gamma, sp.scaled, sp.dt);
    return cudaGetLastError();
}

cudaError_t launch_random_terms(const ParamSets& ps, const double* noise, const int* param_id, double* rnd, int t0, int count,
                                const SimParams& sp, long time0, int rank, long calls_before,
                                const long* calls_dev, int max_ctas, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    RngArgs A{ps, noise, param_id, rnd, t0, count, sp.gamma, sp.scaled, sp.dt, time0, rank, calls_before, calls_dev};
    const int wpb = 4;
    int grid = (count + wpb - 1) / wpb;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    // (max_ctas <= 0: one tetrad per warp.  A balanced resident grid — every warp the same number of tetrads — was
    // measured 19 % slower at 1e5 tetrads: short CTAs handed out by the hardware scheduler beat long-lived warps.)
    random_terms_kernel<<<grid, wpb * 32, 0, st>>>(A);
    return cudaGetLastError();
}

__global__ void bump_counter_kernel(long* counter, long by) { *counter += by; }
cudaError_t launch_bump_counter(long* counter, long by, cudaStream_t st) {
    bump_counter_kernel<<<1, 1, 0, st>>>(counter, by);
    return cudaGetLastError();
}

}  // namespace edmd
