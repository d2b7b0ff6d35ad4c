// nb.cu — inter-tetrad non-bonded forces.
//
// Replaces EDMD::calculate_NB_Forces (/root/reference/src/edmd.cpp:232-285), the per-tetrad
// accumulation of Worker::force_Calculation (src/worker.cpp:165-179, 191-199) and the clip of
// Master::process_NB_Forces (src/master.cpp:297-322).
//
// Two stages:
//
//  distance pass — decides which 32 x 32 atom blocks of which listed tetrad pairs can hold an atom pair
//    that contributes (r^2 below the effective cutoff) and leaves one 64-bit block mask per pair (0 = the
//    pair is not flagged).  Default (edmd_set_nb_scan_mode(4)): exact bounding-sphere culling — group and
//    tetrad spheres (group_bounds_kernel, bounds.cuh), sphere tests per listed pair (pair_select_kernel), then
//    the candidate blocks are evaluated with the biased Gram arithmetic by nb_scan_blocks_kernel (sparse
//    masks: one atom per lane, 48 warps per SM) or nb_scan_cull_kernel (dense masks: 8-atom register tile,
//    partner by one TMA bulk copy).  Brute force (modes 0-2, nb_scan_kernel / nb_scan32_kernel): all
//    num_Atoms x num_Atoms atom pairs of every listed pair, the 9 algorithmic flops per atom pair of
//    edmd.cpp:251-256 in 3 DFMA; this is the kernel the FP64 roofline is defined on.  The
//    cutoff test is an integer min over the high words (r^2 >= 0, so the IEEE bit pattern is monotonic) on
//    the ALU pipe, leaving the FP64 pipe only the distance arithmetic.
//
//  force pass — nb_force_kernel: one warp per (marked tetrad, slot group) re-evaluates the masked blocks
//    of the tetrad's flagged pairs with the reference's exact expression tree (no FMA contraction, IEEE
//    division) and the reference's summation order: per atom over the partner's atoms ascending starting
//    from +0.0, then added to the tetrad total in pair-list order.  Each atom's force is produced by exactly
//    one lane in a fixed order — no atomics, bit-stable, and bit-identical to the W = 1 reference.  Then
//    the clip, and the velocity / coordinate update (finish_*_kernel).
//
// With qfac = 0 (the reference's hard-coded value) an atom pair changes the result only if
// r^2 < 2 (a = max(0, 2 - r^2) > 0): in-cutoff pairs with a = 0 add +-0.0.  The effective cutoff
// is therefore min(atom_Cutoff^2, 2.0) when qfac == 0 and atom_Cutoff^2 otherwise.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "step_math.cuh"
#include "bounds.cuh"

namespace edmd {

__device__ __forceinline__ int hi_word(double v) { return __double2hiint(v); }

// bits g of the result = bits 8g + r of m, g = 0..7 (one multiply gathers the eight strided bits:
// bit 8g' times 2^(56-7g) lands on bit 56+g only for g' = g, and no two products collide)
__device__ __forceinline__ unsigned column_bits(unsigned long long m, int r) {
    return (unsigned)((((m >> r) & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
}

struct ScanArgs {
    const double* crd;      // [n][3][S]
    const int* natoms;      // [n] atoms per tetrad
    const int2* pairs;      // [np] (t1, t2)
    unsigned long long* hit;   // [np] per-pair 8 x 8 block masks, zeroed before the launch; non-zero = the pair is flagged
    long long p0, p1;       // this launch covers pairs [p0, p1)
    int S;
    int hi_cut;             // high word of the effective squared cutoff (MODE 0)
    int chunk;              // pairs per CTA
    double cut;             // effective squared cutoff (MODE 1)
    // Fused compute + collective: when `peers` is set, a hit is stored straight into the hit buffer of
    // EVERY rank (peer pointers over NVLink/NVSwitch, CUDA IPC) instead of being all-gathered later.
    unsigned long long* const* peers;   // [world] device pointers, this rank's own buffer included
    int world;
};

// mask: which 32 x 32 atom blocks of the pair may hold an atom pair inside the cutoff (bit 8g + r: slot
// group r of the lower-numbered tetrad against slot group g of the higher-numbered one); the brute-force
// modes do not track blocks and publish all ones.
__device__ __forceinline__ void publish_hit(const ScanArgs& A, long long k, int lane, unsigned long long mask) {
    if (A.peers) { if (lane < A.world) A.peers[lane][k] = mask; }
    else if (lane == 0) A.hit[k] = mask;
}

constexpr int kScanThreads = 32;                       // one warp = one CTA = one tetrad pair at a time
constexpr int kTI = kMaxStride / kScanThreads;         // 8 register-resident atoms per thread
constexpr int kScanCtasPerSm = 12;

// MODE 0 "direct": r^2 = (xi-xj)^2 + (yi-yj)^2 + (zi-zj)^2 exactly as edmd.cpp:251-256 —
//   3 DADD + DMUL + 2 DFMA per atom pair.
// MODE 1 "gram":   both tetrads are shifted to a pair-local origin o (atom 0 of the register-side
//   tetrad) and the test r^2 < cut is rewritten, with x' = x - o and a bias B = 2^14, as
//       c_ij = (|xj'|^2 + B) - 2 xi'.xj'   <   (cut + B) - |xi'|^2 = thr_i
//   c_ij is a chain of 3 DFMA with (-2 xi') in registers and (xj', |xj'|^2 + B) in shared
//   memory: the same 9 algorithmic flops in 3 FP64-pipe instructions.  c_ij = r^2 - |xi'|^2 + B
//   is positive whenever |xi'|^2 < B (|xi'| < 128 A; a thread that sees a larger value flags
//   the pair unconditionally), so the high-word integer test applies: hi(c) <= hi(thr_i) + 1.
//   Rounding error of the chain <= 2^-51 (|xi'|+|xj'|)^2 + 2^-51 B ~ 1e-11, slack of the test
//   >= 2^-20 thr_i >= 2e-6: a superset of the true hits, always.
//   Either way the scan only SELECTS pairs: the forces come from nb_force_kernel's exact
//   arithmetic, so both modes give bit-identical results.
//
// Instruction-issue budget (measured on B200: an FP64 warp instruction holds the dispatch port
// for 2 cycles): per atom pair MODE 0 issues 6 FP64 + 0.375 LDS.128 + 0.5 VIMNMX3, MODE 1 issues
// 3 FP64 + 0.5 LDS.128 + 0.5 VIMNMX3 — the 4-atom register tile is what keeps the non-FP64 share
// small (the first version of this kernel, 1 atom per thread, was dispatch-bound at 80 % FP64).
// (Tried and rejected: one partner atom per un-unrolled iteration, which makes ptxas emit 8-long runs
// of DFMAs sharing the partner operand — full-rate register-file pattern — but costs more in loop
// overhead and exposed LDS latency than it gains: 8.03 vs 7.37 ms on C1.)
// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a) ------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one contiguous global->shared bulk copy (UBLKCP), completion signalled on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes,
                                            unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kScanThreads, kScanCtasPerSm) nb_scan_kernel(ScanArgs A) {
    constexpr int NP = MODE == 1 ? 4 : 3;                // planes staged per tetrad
    constexpr double kBias = 16384.0;                    // B, see above
    constexpr int NB = kScanThreads == 32 ? 1 : 2;       // a single warp needs no double buffer
    __shared__ __align__(16) double sb[NB][NP][kMaxStride + 2];  // staged partner (+2: prefetch overrun)
    __shared__ __align__(128) double raw[3 * kMaxStride];    // TMA landing zone: next partner's x|y|z planes
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, lane = tid & 31;
    const int S = A.S;

    // a short contiguous chunk of the pair list per CTA (the register-side tetrad — the lower
    // index, i.e. the row of the i-major list — rarely changes inside a chunk); many more CTAs
    // than SM slots, so the hardware block scheduler balances the tail dynamically
    const long long chunk = A.chunk;
    const long long k0 = A.p0 + chunk * blockIdx.x;
    const long long k1 = (k0 + chunk < A.p1) ? k0 + chunk : A.p1;
    if (k0 >= k1) return;
    const unsigned plane_bytes = 3u * (unsigned)S * 8u;  // one tetrad = one contiguous block

    double xi[kTI], yi[kTI], zi[kTI];                    // MODE 1: -2x', -2y', -2z'
    int hthr[kTI];                                       // MODE 1: hi(thr_i) + 1, or -1 for a padding atom
    double ox = 0.0, oy = 0.0, oz = 0.0;
    int cur_r = -1;
    bool wild = false;                                   // MODE 1: |xi'| too large for the error bound

    int2 pr = A.pairs[k0];
    if (tid == 0) {
        mbar_init(&bar, 1);
        const int ts0 = pr.x < pr.y ? pr.y : pr.x;
        tma_load_1d(raw, A.crd + (size_t)ts0 * 3 * S, plane_bytes, &bar);
    }
    if (kScanThreads == 32) __syncwarp(); else __syncthreads();   // mbarrier initialised before anyone waits

    int b = 0;
    unsigned phase = 0;
    for (long long k = k0; k < k1; k++, b ^= (NB - 1), phase ^= 1) {
        const int tr = pr.x < pr.y ? pr.x : pr.y;
        const int ns = A.natoms[pr.x < pr.y ? pr.y : pr.x];
        int2 pr_next = pr;
        if (k + 1 < k1) pr_next = A.pairs[k + 1];
        if (tr != cur_r) {                               // new row: reload the register tile
            cur_r = tr;
            const double* Xr = A.crd + (size_t)tr * 3 * S;
            const int nr = A.natoms[tr];
            if (MODE == 1) { ox = Xr[0]; oy = Xr[S]; oz = Xr[2 * S]; wild = false; }
#pragma unroll
            for (int r = 0; r < kTI; r++) {
                const int a = tid + kScanThreads * r;
                if (a < nr) {
                    const double x = Xr[a], y = Xr[S + a], z = Xr[2 * S + a];
                    if (MODE == 1) {
                        const double lx = x - ox, ly = y - oy, lz = z - oz;
                        xi[r] = -2.0 * lx; yi[r] = -2.0 * ly; zi[r] = -2.0 * lz;
                        const double w = fma(lz, lz, fma(ly, ly, lx * lx));
                        wild = wild || !(w < 0.99 * kBias);
                        hthr[r] = hi_word((A.cut + kBias) - w) + 1;
                    } else { xi[r] = x; yi[r] = y; zi[r] = z; }
                } else if (MODE == 1) { xi[r] = yi[r] = zi[r] = 0.0; hthr[r] = -1; }
                else { xi[r] = yi[r] = zi[r] = -1e150; }
            }
        }
        // stage the partner that the TMA engine delivered during the previous pair's arithmetic
        mbar_wait(&bar, phase);
        const int ns2 = (ns + 1) & ~1;
#pragma unroll
        for (int r = 0; r < kTI; r++) {
            const int a = tid + kScanThreads * r;
            if (a < ns2) {
                const bool real = a < ns;
                const double rx = raw[a], ry = raw[S + a], rz = raw[2 * S + a];
                if (MODE == 1) {
                    const double lx = rx - ox, ly = ry - oy, lz = rz - oz;
                    sb[b][0][a] = real ? lx : 0.0; sb[b][1][a] = real ? ly : 0.0; sb[b][2][a] = real ? lz : 0.0;
                    sb[b][3][a] = real ? fma(lz, lz, fma(ly, ly, lx * lx)) + kBias : 1e300;
                } else {
                    sb[b][0][a] = real ? rx : 1e150; sb[b][1][a] = real ? ry : 1e150; sb[b][2][a] = real ? rz : 1e150;
                }
            }
        }
        if (kScanThreads == 32) __syncwarp(); else __syncthreads();   // staged[b] complete, raw consumed
        if (tid == 0 && k + 1 < k1) {                    // next partner lands while this pair is computed
            const int tsn = pr_next.x < pr_next.y ? pr_next.y : pr_next.x;
            tma_load_1d(raw, A.crd + (size_t)tsn * 3 * S, plane_bytes, &bar);
        }

        int mm[kTI];
#pragma unroll
        for (int r = 0; r < kTI; r++) mm[r] = 0x7fffffff;
        const double* bx = sb[b][0]; const double* by = sb[b][1]; const double* bz = sb[b][2];
        const double* bw = sb[b][NP - 1];
        // software-pipelined: operands of step j+2 are fetched before step j's arithmetic
        double2 xs = *reinterpret_cast<const double2*>(bx), ys = *reinterpret_cast<const double2*>(by);
        double2 zs = *reinterpret_cast<const double2*>(bz), ws = *reinterpret_cast<const double2*>(bw);
#pragma unroll 4
        for (int j = 0; j < ns2; j += 2) {
            const double2 nxs = *reinterpret_cast<const double2*>(bx + j + 2);
            const double2 nys = *reinterpret_cast<const double2*>(by + j + 2);
            const double2 nzs = *reinterpret_cast<const double2*>(bz + j + 2);
            double2 nws = ws;
            if (MODE == 1) nws = *reinterpret_cast<const double2*>(bw + j + 2);
            if (MODE == 1) {
                // r innermost at every chain level: consecutive DFMAs share the shared-memory
                // operand (operand-reuse cache), so each reads only 2 registers from the RF — a DFMA
                // with 3 distinct register sources is limited to one issue per 3 cycles
                double c0[kTI], c1[kTI];
#pragma unroll
                for (int r = 0; r < kTI; r++) c0[r] = fma(zi[r], zs.x, ws.x);
#pragma unroll
                for (int r = 0; r < kTI; r++) c0[r] = fma(yi[r], ys.x, c0[r]);
#pragma unroll
                for (int r = 0; r < kTI; r++) c0[r] = fma(xi[r], xs.x, c0[r]);
#pragma unroll
                for (int r = 0; r < kTI; r++) c1[r] = fma(zi[r], zs.y, ws.y);
#pragma unroll
                for (int r = 0; r < kTI; r++) c1[r] = fma(yi[r], ys.y, c1[r]);
#pragma unroll
                for (int r = 0; r < kTI; r++) c1[r] = fma(xi[r], xs.y, c1[r]);
#pragma unroll
                for (int r = 0; r < kTI; r++) mm[r] = min(mm[r], min(hi_word(c0[r]), hi_word(c1[r])));
            } else {
#pragma unroll
                for (int r = 0; r < kTI; r++) {
                    const double dx0 = xi[r] - xs.x, dy0 = yi[r] - ys.x, dz0 = zi[r] - zs.x;
                    const double dx1 = xi[r] - xs.y, dy1 = yi[r] - ys.y, dz1 = zi[r] - zs.y;
                    const double r0 = fma(dz0, dz0, fma(dy0, dy0, dx0 * dx0));
                    const double r1 = fma(dz1, dz1, fma(dy1, dy1, dx1 * dx1));
                    mm[r & 1] = min(mm[r & 1], min(hi_word(r0), hi_word(r1)));
                }
            }
            xs = nxs; ys = nys; zs = nzs; ws = nws;
        }
        bool pred = false;
        if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < kTI; r++) pred = pred || (mm[r] <= hthr[r]);
            pred = pred || wild;
        } else {
            pred = min(mm[0], mm[1]) <= A.hi_cut;
        }
        if (__any_sync(0xffffffffu, pred)) publish_hit(A, k, lane, ~0ull);   // (the vote is also the WAR fence of the single buffer)
        pr = pr_next;
    }
}

// ---- MODE 4: gram form + bounding-sphere culling ---------------------------------------------------
// The 256 slots of a tetrad are filled with its atoms through a per-parameter-set permutation that
// makes slots 32g .. 32g+31 a spatially compact group g (recursive coordinate bisection of the
// average structure, engine.cu); in the register tile group r is exactly register r.
//   group_bounds_kernel   a bounding sphere for every group and for every whole tetrad that is an
//                         end of one of this engine's pairs (mark_needed_kernel);
//   pair_select_kernel    one tetrad-sphere test per listed pair, then the 8 x 8 group-sphere tests
//                         -> a 64-bit mask (bit 8g + r: register group r may be within sqrt(cut) of
//                         partner group g); pairs with a non-empty mask go to a compact work list;
//   nb_scan_cull_kernel   walks the work list: TMA-loads the partner, stages it in group order and
//                         evaluates only the masked 32 x 32 blocks with the gram arithmetic of MODE 1.
// Conservative by construction (spheres contain their atoms, every test is inflated), so the selected
// pairs remain a superset of the true hits and the final forces are bit-identical to modes 0-2
// (tests/test_gpu_edge.py).  This is the engine's default; modes 0-2 evaluate all 63,504 atom pairs
// of every listed pair like the reference does and are kept for the roofline measurement of the
// distance kernel (bench.py `variants`).
__global__ void mark_needed_kernel(const int2* pairs, long long p0, long long p1, unsigned char* need) {
    const long long k = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p1) return;
    const int2 pr = pairs[k];
    need[pr.x] = 1; need[pr.y] = 1;
}
__global__ void need_compact_kernel(const unsigned char* need, int n, int* list, int* count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && need[t]) list[atomicAdd(count, 1)] = t;
}
// need[t] = 1 for both ends of the pairs in [p0, p1); list / count: the same tetrads as a compact list (any order)
cudaError_t launch_mark_needed(const int2* pairs, long long p0, long long p1, unsigned char* need, int* list, int* count,
                               int n, cudaStream_t st) {
    cudaError_t err = cudaMemsetAsync(need, 0, (size_t)n, st);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(count, 0, sizeof(int), st);
    if (err != cudaSuccess || p1 <= p0) return err;
    mark_needed_kernel<<<(unsigned)((p1 - p0 + 255) / 256), 256, 0, st>>>(pairs, p0, p1, need);
    need_compact_kernel<<<(n + 255) / 256, 256, 0, st>>>(need, n, list, count);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(128, 10) group_bounds_kernel(const double* crd, const int* param_id, const int* perm,
                                                            double* gb, int n, int S, const int* need_list, int need_count) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= need_count) return;
    const int t = need_list[i];  // an end of some pair in this engine's scan slice (sharded runs: a fraction of all)
    group_bounds_tetrad(crd + (size_t)t * 3 * S, perm + (size_t)param_id[t] * kMaxStride, gb, n, t, S, lane);
}

// One thread per listed pair: whole-tetrad spheres first (one test instead of 64 for the typical
// far-apart pair), then the 8 x 8 group spheres.  Pairs with a non-empty mask are appended to a
// compact work list {pair index, mask} (warp-aggregated atomic; the order of the list varies from
// run to run, the set — and therefore every hit flag — does not).
struct CullItem { long long k; unsigned long long mask; };

__global__ void __launch_bounds__(128) pair_select_kernel(const double* gb, const int2* pairs, long long p0, long long p1,
                                                          double dcut, CullItem* items, CullItem* items_back, unsigned* count,
                                                          int dense_min, int n) {
    const long long k = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long m = 0;
    if (k < p1) {
        const int2 pr = pairs[k];
        const int tr = pr.x < pr.y ? pr.x : pr.y, ts = pr.x < pr.y ? pr.y : pr.x;
        const double4* TB = reinterpret_cast<const double4*>(gb) + (size_t)n * 8;
        const double4 a = TB[tr], b = TB[ts];
        const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
        const double lim = a.w + b.w + dcut;
        if (a.w >= 0.0 && b.w >= 0.0 && !(dx * dx + dy * dy + dz * dz > lim * lim * (1.0 + 1e-9))) {
            const double4* R = reinterpret_cast<const double4*>(gb) + (size_t)tr * 8;
            const double4* P = reinterpret_cast<const double4*>(gb) + (size_t)ts * 8;
            double4 rr[8];
#pragma unroll
            for (int r = 0; r < 8; r++) rr[r] = R[r];
#pragma unroll 1
            for (int g = 0; g < 8; g++) {
                const double4 q = P[g];
                if (q.w < 0.0) continue;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const double ex = rr[r].x - q.x, ey = rr[r].y - q.y, ez = rr[r].z - q.z;
                    const double l2 = rr[r].w + q.w + dcut;
                    if (rr[r].w >= 0.0 && ex * ex + ey * ey + ez * ez <= l2 * l2 * (1.0 + 1e-9)) m |= 1ull << (8 * g + r);
                }
            }
        }
    }
    // sparse masks go to the front list (count[0], ascending from items), dense ones to the back list (count[1],
    // descending from items_back): one buffer of one item per listed pair serves both
    const int lane = threadIdx.x & 31;
    const bool dense = m != 0ull && __popcll(m) >= dense_min;
    const unsigned live = __ballot_sync(0xffffffffu, m != 0ull && !dense);
    const unsigned lived = __ballot_sync(0xffffffffu, dense);
    if (live) {
        unsigned base = 0;
        if (lane == __ffs(live) - 1) base = atomicAdd(count, (unsigned)__popc(live));
        base = __shfl_sync(0xffffffffu, base, __ffs(live) - 1);
        if (m && !dense) items[base + __popc(live & ((1u << lane) - 1u))] = CullItem{k, m};
    }
    if (lived) {
        unsigned base = 0;
        if (lane == __ffs(lived) - 1) base = atomicAdd(count + 1, (unsigned)__popc(lived));
        base = __shfl_sync(0xffffffffu, base, __ffs(lived) - 1);
        if (dense) *(items_back - (long long)(base + __popc(lived & ((1u << lane) - 1u)))) = CullItem{k, m};
    }
}

__global__ void __launch_bounds__(kScanThreads, 15) nb_scan_cull_kernel(ScanArgs A, const CullItem* items_back,
                                                                                    const unsigned* count,
                                                                                    const int* param_id, const int* perm) {
    // items_back: the dense-mask list, item i at items_back[-i] (pair_select_kernel)
    constexpr double kBias = 16384.0;
    __shared__ __align__(16) double sb[4][kMaxStride + 2];
    __shared__ __align__(128) double raw[3 * kMaxStride];
    __shared__ __align__(8) unsigned long long bar;
    const int lane = threadIdx.x & 31;
    const int S = A.S;
    const unsigned plane_bytes = 3u * (unsigned)S * 8u;
    // The work list is cut into blocks of `chunk` items (one per lane, chunk <= 32), chunk chosen from
    // the DEVICE-side item count so that a short list still spreads over every SM (one item per warp)
    // and a long one amortises the register-side tetrad over a run of its partners.
    const unsigned total = *count;
    unsigned chunk = (total + (unsigned)A.chunk - 1) / (unsigned)A.chunk;     // A.chunk = blocks wanted (64 per SM)
    chunk = chunk < 1 ? 1 : (chunk > 32 ? 32 : chunk);
    if ((unsigned long long)blockIdx.x * chunk >= total) return;

    double xi[kTI], yi[kTI], zi[kTI];
    int hthr[kTI];
    double ox = 0.0, oy = 0.0, oz = 0.0;
    int cur_r = -1;
    bool wild = false;
    if (lane == 0) mbar_init(&bar, 1);
    __syncwarp();
    unsigned phase = 0;
    for (unsigned long long blk = blockIdx.x; blk * chunk < total; blk += gridDim.x) {
    const unsigned i0 = (unsigned)(blk * chunk);
    const unsigned i1 = (i0 + chunk < total) ? i0 + chunk : total;
    CullItem mine{0, 0ull};
    if (i0 + lane < i1) mine = *(items_back - (long long)(i0 + lane));
    const int nitems = (int)(i1 - i0);
    int it = 0;
    long long k = __shfl_sync(0xffffffffu, mine.k, 0);
    int2 pr = A.pairs[k];
    if (lane == 0) tma_load_1d(raw, A.crd + (size_t)(pr.x < pr.y ? pr.y : pr.x) * 3 * S, plane_bytes, &bar);
    while (it < nitems) {
        const unsigned long long m = __shfl_sync(0xffffffffu, mine.mask, it);
        const bool more = it + 1 < nitems;
        const long long kn = __shfl_sync(0xffffffffu, mine.k, more ? it + 1 : it);
        int2 pr_next = pr;
        if (more) pr_next = A.pairs[kn];
        const int tr = pr.x < pr.y ? pr.x : pr.y;
        const int tsc = pr.x < pr.y ? pr.y : pr.x;
        if (tr != cur_r) {
            cur_r = tr;
            const double* Xr = A.crd + (size_t)tr * 3 * S;
            const int* PR = perm + (size_t)param_id[tr] * kMaxStride;
            ox = Xr[0]; oy = Xr[S]; oz = Xr[2 * S]; wild = false;
#pragma unroll
            for (int r = 0; r < kTI; r++) {
                const int a = PR[lane + kScanThreads * r];     // slot -> atom
                if (a >= 0) {
                    const double lx = Xr[a] - ox, ly = Xr[S + a] - oy, lz = Xr[2 * S + a] - oz;
                    xi[r] = -2.0 * lx; yi[r] = -2.0 * ly; zi[r] = -2.0 * lz;
                    const double w = fma(lz, lz, fma(ly, ly, lx * lx));
                    wild = wild || !(w < 0.99 * kBias);
                    hthr[r] = hi_word((A.cut + kBias) - w) + 1;
                } else { xi[r] = yi[r] = zi[r] = 0.0; hthr[r] = -1; }
            }
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        {   // stage the partner's slot groups that some masked block needs, in compact-group order (empty slots never hit)
            const int* PS = perm + (size_t)param_id[tsc] * kMaxStride;
#pragma unroll
            for (int r = 0; r < kTI; r++) {
                if (!((m >> (8 * r)) & 0xffull)) continue;      // partner group r takes part in no masked block
                const int slot = lane + kScanThreads * r;
                const int a = PS[slot];
                const bool real = a >= 0;
                const double lx = real ? raw[a] - ox : 0.0, ly = real ? raw[S + a] - oy : 0.0, lz = real ? raw[2 * S + a] - oz : 0.0;
                sb[0][slot] = lx; sb[1][slot] = ly; sb[2][slot] = lz;
                sb[3][slot] = real ? fma(lz, lz, fma(ly, ly, lx * lx)) + kBias : 1e300;
            }
        }
        __syncwarp();
        if (lane == 0 && more) {
            const int tsn = pr_next.x < pr_next.y ? pr_next.y : pr_next.x;
            tma_load_1d(raw, A.crd + (size_t)tsn * 3 * S, plane_bytes, &bar);
        }
        // Evaluate the masked blocks; after each partner group g the per-lane results are folded into
        // ONE warp-wide 8-bit OR (redux.sync): bit r = some atom of my group r lies within the cutoff of
        // some atom of partner group g.  The refined mask goes to the force pass, which then touches
        // only those 32 x 32 blocks.
        unsigned long long refined = 0ull;
#pragma unroll 1
        for (int g = 0; g < 8; g++) {
            const unsigned m8 = (unsigned)(m >> (8 * g)) & 0xffu;
            if (!m8) continue;
            int mm[kTI];
#pragma unroll
            for (int r = 0; r < kTI; r++) mm[r] = 0x7fffffff;
            const int jhi = 32 * g + 32;
#pragma unroll 2
            for (int j = 32 * g; j < jhi; j += 2) {
                const double2 xs = *reinterpret_cast<const double2*>(sb[0] + j);
                const double2 ys = *reinterpret_cast<const double2*>(sb[1] + j);
                const double2 zs = *reinterpret_cast<const double2*>(sb[2] + j);
                const double2 ws = *reinterpret_cast<const double2*>(sb[3] + j);
#pragma unroll
                for (int r = 0; r < kTI; r++) {
                    if (m8 & (1u << r)) {
                        const double c0 = fma(xi[r], xs.x, fma(yi[r], ys.x, fma(zi[r], zs.x, ws.x)));
                        const double c1 = fma(xi[r], xs.y, fma(yi[r], ys.y, fma(zi[r], zs.y, ws.y)));
                        mm[r] = min(mm[r], min(hi_word(c0), hi_word(c1)));
                    }
                }
            }
            unsigned mine8 = 0;
#pragma unroll
            for (int r = 0; r < kTI; r++) mine8 |= (mm[r] <= hthr[r]) ? (1u << r) : 0u;
            refined |= (unsigned long long)__reduce_or_sync(0xffffffffu, mine8) << (8 * g);
        }
        // a tetrad too spread out for the bias: keep every block the sphere tests let through
        if (__any_sync(0xffffffffu, wild)) refined = m;
        if (refined) publish_hit(A, k, lane, refined);
        pr = pr_next;
        k = kn;
        it++;
    }
    }   // blocks of the work list
}

// Light-weight variant of the masked distance kernel for SPARSE masks (the usual case: 2-3 of a pair's 64 blocks
// survive the sphere tests).  nb_scan_cull_kernel above keeps a whole tetrad in an 8-atom register tile and
// pulls the whole partner by TMA — 122 registers, 14 KB of shared memory, 15 warps per SM — and spends ~9 us per
// item on latencies it cannot overlap, whatever the number of blocks.  Here one warp takes one item and walks its
// set bits; a lane holds ONE atom of the current group of the lower-numbered tetrad, the 32 atoms of the partner
// group are gathered through the slot table into a 1 KB shared strip and read as broadcasts: ~40 registers, so
// the SM holds 48 warps and the per-item latency is hidden by other items.  Same Gram arithmetic, same
// conservative threshold, same refined mask; dense masks (dense_min or more blocks per pair) still go to the tiled
// kernel, which amortises the partner loads over eight groups (pair_select_kernel files the items in two lists).
constexpr int kBlkWarps = 4;

__global__ void __launch_bounds__(32 * kBlkWarps, 12) nb_scan_blocks_kernel(ScanArgs A, const CullItem* items,
                                                                            const unsigned* count,
                                                                            const int* param_id, const int* perm) {
    constexpr double kBias = 16384.0;
    __shared__ __align__(16) double4 sp[kBlkWarps][32];      // partner group: x', y', z', |x'|^2 + B
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wslot = threadIdx.x >> 5;
    const int S = A.S;
    const unsigned total = __reduce_max_sync(full, *count);
    const unsigned w0 = __reduce_max_sync(full, blockIdx.x * kBlkWarps + wslot);
    for (unsigned item = w0; item < total; item += gridDim.x * kBlkWarps) {
        const CullItem it = items[item];
        const unsigned mlo = __reduce_max_sync(full, (unsigned)it.mask), mhi = __reduce_max_sync(full, (unsigned)(it.mask >> 32));
        const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
        const int k = (int)__reduce_max_sync(full, (unsigned)it.k);
        const int2 pr = A.pairs[k];
        const int tr = __reduce_min_sync(full, pr.x < pr.y ? pr.x : pr.y), ts = __reduce_max_sync(full, pr.x < pr.y ? pr.y : pr.x);
        const double* Xr = A.crd + (size_t)tr * 3 * S;
        const double* Xs = A.crd + (size_t)ts * 3 * S;
        const int* PR = perm + (size_t)param_id[tr] * kMaxStride;
        const int* PS = perm + (size_t)param_id[ts] * kMaxStride;
        const double ox = Xr[0], oy = Xr[S], oz = Xr[2 * S];   // pair-local origin, as in the tiled kernel
        unsigned long long refined = 0ull;
        bool wild = false;
#pragma unroll 1
        for (int r = 0; r < 8; r++) {
            const unsigned mr = column_bits(m, r);             // partner groups that may reach my group r
            if (!mr) continue;
            const int a = PR[32 * r + lane];
            double xi = 0.0, yi = 0.0, zi = 0.0; int hthr = -1;
            if (a >= 0) {
                const double lx = Xr[a] - ox, ly = Xr[S + a] - oy, lz = Xr[2 * S + a] - oz;
                xi = -2.0 * lx; yi = -2.0 * ly; zi = -2.0 * lz;
                const double w = fma(lz, lz, fma(ly, ly, lx * lx));
                wild = wild || !(w < 0.99 * kBias);
                hthr = hi_word((A.cut + kBias) - w) + 1;
            }
            unsigned todo = mr;
            while (todo) {
                const int g = __ffs(todo) - 1;
                todo &= todo - 1;
                const int b = PS[32 * g + lane];
                double4 v = make_double4(0.0, 0.0, 0.0, 1e300);   // an empty slot never hits
                if (b >= 0) {
                    const double lx = Xs[b] - ox, ly = Xs[S + b] - oy, lz = Xs[2 * S + b] - oz;
                    v = make_double4(lx, ly, lz, fma(lz, lz, fma(ly, ly, lx * lx)) + kBias);
                }
                sp[wslot][lane] = v;
                __syncwarp();
                int mm = 0x7fffffff;
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    const double4 q = sp[wslot][j];
                    mm = min(mm, hi_word(fma(xi, q.x, fma(yi, q.y, fma(zi, q.z, q.w)))));
                }
                if (__any_sync(full, mm <= hthr)) refined |= 1ull << (8 * g + r);
                __syncwarp();
            }
        }
        // a tetrad too spread out for the bias: keep every block the sphere tests let through
        if (__any_sync(full, wild)) refined = m;
        if (refined) publish_hit(A, k, lane, refined);
    }
}

cudaError_t launch_nb_scan_cull(const double* crd, const int* natoms, const int2* pairs, unsigned long long* hit,
                                long long p0, long long p1, int n, int S, double cut2_eff, int sms, double* gb,
                                unsigned long long* work, long long work_items, unsigned* count, int dense_min,
                                const int* param_id, const int* perm, const int* need_list, int need_count,
                                unsigned long long* const* peers, int world, cudaStream_t st) {
    if (p1 <= p0) return cudaSuccess;
    if (!peers) {
        cudaError_t err = cudaMemsetAsync(hit + p0, 0, sizeof(unsigned long long) * (size_t)(p1 - p0), st);
        if (err != cudaSuccess) return err;
    }
    // work buffer: one CullItem per listed pair at most; the two item counters were zeroed by the caller
    CullItem* items = reinterpret_cast<CullItem*>(work);
    CullItem* items_back = items + (work_items - 1);
    if (need_count > 0)
        group_bounds_kernel<<<(need_count + 3) / 4, 128, 0, st>>>(crd, param_id, perm, gb, n, S, need_list, need_count);
    const long long want = p1 - p0;
    pair_select_kernel<<<(unsigned)((want + 127) / 128), 128, 0, st>>>(gb, pairs, p0, p1, sqrt(cut2_eff), items, items_back, count,
                                                                     dense_min, n);
    // The number of selected pairs is only known on the device: fixed grids walk the lists; CTAs beyond a list
    // read its counter and exit.  pair_select_kernel files an item by the population of its sphere mask: sparse
    // masks (< dense_min blocks) at the front of the work buffer for the light-weight block kernel, dense ones at the
    // back for the tiled kernel (count[0] / count[1]; the two lists grow towards each other).
    const long long blocks_wanted = (long long)sms * 64;
    long long bits; memcpy(&bits, &cut2_eff, 8);
    ScanArgs A{crd, natoms, pairs, hit, p0, p1, S, (int)(bits >> 32), (int)blocks_wanted, cut2_eff, peers, world};
    if (dense_min > 1) {          // (dense_min <= 1: every item is "dense", tiled kernel only)
        const long long warps = want < 48LL * sms ? want : 48LL * sms;
        nb_scan_blocks_kernel<<<(int)((warps + kBlkWarps - 1) / kBlkWarps), 32 * kBlkWarps, 0, st>>>(A, items, count, param_id, perm);
    }
    if (dense_min <= 64) {        // (dense_min > 64: no mask is dense, block kernel only)
        // (its CTAs walk the list with a grid stride.  With the steric cutoff alone (sqrt(2) A) a pair rarely has nine
        // candidate blocks, and an almost idle grid of 4,440 one-warp CTAs would cost 16 us per step: a small grid then,
        // two residents' worth when the cutoff is the electrostatic one)
        const long long resident = (cut2_eff > 4.0 ? 30LL : 4LL) * sms;
        const int grid = (int)(want < resident ? want : resident);
        nb_scan_cull_kernel<<<grid, kScanThreads, 0, st>>>(A, items_back, count + 1, param_id, perm);
    }
    return cudaGetLastError();
}

// ---- MODE 2: FP32 prefilter with packed FFMA2 -----------------------------------------------------
// The "validated FP32 variant" of the distance pass (BASELINE north_star, configs[4]): the same biased
// Gram test as MODE 1 evaluated in single precision, two partner atoms per instruction
// (fma.rn.f32x2, SASS FFMA2 — Blackwell's packed FP32 pipe).  It stays a SUPERSET filter:
//   B = 4096 (|xi'| < 64 A, else the pair is flagged unconditionally), cut <= 100;
//   inputs rounded to float:       |dc| <= 2*sqrt(3)*(64*2^-24*74 + 74*2^-24*64) < 0.002
//   |xj'|^2 + B rounded once from FP64:                                   <= 0.0005
//   three FMA roundings at magnitude < 2^15:                              <= 3 * 2^-9 = 0.006
// so |c_f32 - c| < 0.01 for every atom pair that can be a hit, and the threshold carries a margin
// of 0.02.  Forces still come from nb_accumulate's exact FP64 path: results are bit-identical to
// the FP64 modes (tests/test_gpu_edge.py).
__global__ void __launch_bounds__(kScanThreads, 16) nb_scan32_kernel(ScanArgs A) {
    constexpr float kBias = 4096.0f;
    constexpr int kRow = kMaxStride + 8;
    __shared__ __align__(16) float sb[4][kRow];
    __shared__ __align__(128) double raw[3 * kMaxStride];
    __shared__ __align__(8) unsigned long long bar;
    const int lane = threadIdx.x & 31;
    const int S = A.S;
    const long long chunk = A.chunk;
    const long long k0 = A.p0 + chunk * blockIdx.x;
    const long long k1 = (k0 + chunk < A.p1) ? k0 + chunk : A.p1;
    if (k0 >= k1) return;
    const unsigned plane_bytes = 3u * (unsigned)S * 8u;

    float2 xi[kTI], yi[kTI], zi[kTI];
    int hthr[kTI];
    double ox = 0.0, oy = 0.0, oz = 0.0;
    int cur_r = -1;
    bool wild = false;

    int2 pr = A.pairs[k0];
    if (lane == 0) {
        mbar_init(&bar, 1);
        const int ts0 = pr.x < pr.y ? pr.y : pr.x;
        tma_load_1d(raw, A.crd + (size_t)ts0 * 3 * S, plane_bytes, &bar);
    }
    __syncwarp();
    unsigned phase = 0;
    for (long long k = k0; k < k1; k++, phase ^= 1) {
        const int tr = pr.x < pr.y ? pr.x : pr.y;
        const int ns = A.natoms[pr.x < pr.y ? pr.y : pr.x];
        int2 pr_next = pr;
        if (k + 1 < k1) pr_next = A.pairs[k + 1];
        if (tr != cur_r) {
            cur_r = tr;
            const double* Xr = A.crd + (size_t)tr * 3 * S;
            const int nr = A.natoms[tr];
            ox = Xr[0]; oy = Xr[S]; oz = Xr[2 * S]; wild = false;
#pragma unroll
            for (int r = 0; r < kTI; r++) {
                const int a = lane + kScanThreads * r;
                if (a < nr) {
                    const double lx = Xr[a] - ox, ly = Xr[S + a] - oy, lz = Xr[2 * S + a] - oz;
                    const float fx = (float)(-2.0 * lx), fy = (float)(-2.0 * ly), fz = (float)(-2.0 * lz);
                    xi[r] = make_float2(fx, fx); yi[r] = make_float2(fy, fy); zi[r] = make_float2(fz, fz);
                    const double w = fma(lz, lz, fma(ly, ly, lx * lx));
                    wild = wild || !(w < 0.99 * (double)kBias);
                    hthr[r] = __float_as_int(__double2float_ru((A.cut + (double)kBias) - w + 0.02)) + 1;
                } else {
                    xi[r] = yi[r] = zi[r] = make_float2(0.f, 0.f); hthr[r] = -1;
                }
            }
        }
        mbar_wait(&bar, phase);
        const int ns4 = (ns + 3) & ~3;
#pragma unroll
        for (int r = 0; r < kTI; r++) {
            const int a = lane + kScanThreads * r;
            if (a < ns4) {
                const bool real = a < ns;
                const double lx = raw[a] - ox, ly = raw[S + a] - oy, lz = raw[2 * S + a] - oz;
                sb[0][a] = real ? (float)lx : 0.f; sb[1][a] = real ? (float)ly : 0.f; sb[2][a] = real ? (float)lz : 0.f;
                sb[3][a] = real ? (float)(fma(lz, lz, fma(ly, ly, lx * lx)) + (double)kBias) : 1e30f;
            }
        }
        __syncwarp();
        if (lane == 0 && k + 1 < k1) {
            const int tsn = pr_next.x < pr_next.y ? pr_next.y : pr_next.x;
            tma_load_1d(raw, A.crd + (size_t)tsn * 3 * S, plane_bytes, &bar);
        }
        int mm[kTI];
#pragma unroll
        for (int r = 0; r < kTI; r++) mm[r] = 0x7fffffff;
        float4 xs = *reinterpret_cast<const float4*>(sb[0]), ys = *reinterpret_cast<const float4*>(sb[1]);
        float4 zs = *reinterpret_cast<const float4*>(sb[2]), ws = *reinterpret_cast<const float4*>(sb[3]);
#pragma unroll 4
        for (int j = 0; j < ns4; j += 4) {
            const float4 nxs = *reinterpret_cast<const float4*>(sb[0] + j + 4);
            const float4 nys = *reinterpret_cast<const float4*>(sb[1] + j + 4);
            const float4 nzs = *reinterpret_cast<const float4*>(sb[2] + j + 4);
            const float4 nws = *reinterpret_cast<const float4*>(sb[3] + j + 4);
            const float2 xa = make_float2(xs.x, xs.y), xb = make_float2(xs.z, xs.w);
            const float2 ya = make_float2(ys.x, ys.y), yb = make_float2(ys.z, ys.w);
            const float2 za = make_float2(zs.x, zs.y), zb = make_float2(zs.z, zs.w);
            const float2 wa = make_float2(ws.x, ws.y), wb = make_float2(ws.z, ws.w);
#pragma unroll
            for (int r = 0; r < kTI; r++) {
                const float2 ca = __ffma2_rn(xi[r], xa, __ffma2_rn(yi[r], ya, __ffma2_rn(zi[r], za, wa)));
                const float2 cb = __ffma2_rn(xi[r], xb, __ffma2_rn(yi[r], yb, __ffma2_rn(zi[r], zb, wb)));
                mm[r] = min(mm[r], min(__float_as_int(ca.x), __float_as_int(ca.y)));
                mm[r] = min(mm[r], min(__float_as_int(cb.x), __float_as_int(cb.y)));
            }
            xs = nxs; ys = nys; zs = nzs; ws = nws;
        }
        bool pred = wild;
#pragma unroll
        for (int r = 0; r < kTI; r++) pred = pred || (mm[r] <= hthr[r]);
        if (__any_sync(0xffffffffu, pred)) publish_hit(A, k, lane, ~0ull);
        pr = pr_next;
    }
}

cudaError_t launch_nb_scan(const double* crd, const int* natoms, const int2* pairs, unsigned long long* hit,
                           long long p0, long long p1, int S, double cut2_eff, int mode, int sms,
                           unsigned long long* const* peers, int world, cudaStream_t st) {
    if (p1 <= p0) return cudaSuccess;
    if (!peers) {   // (peer mode: the buffers are zeroed by the double-buffer protocol in engine.cu)
        cudaError_t err = cudaMemsetAsync(hit + p0, 0, sizeof(unsigned long long) * (size_t)(p1 - p0), st);
        if (err != cudaSuccess) return err;
    }
    const long long want = p1 - p0;
    long long chunk = want / ((long long)sms * 64);      // >= 64 chunks per SM when there is enough work
    if (chunk < 1) chunk = 1;
    if (chunk > 8) chunk = 8;
    const int grid = (int)((want + chunk - 1) / chunk);
    // hi-word test: r^2 < cut  ==>  hi(r^2) <= hi(cut); a superset, re-checked exactly later.
    long long bits; memcpy(&bits, &cut2_eff, 8);
    const bool gram = (mode >= 1) && cut2_eff >= 1e-3 && cut2_eff < 1e6;
    ScanArgs A{crd, natoms, pairs, hit, p0, p1, S, (int)(bits >> 32), (int)chunk, cut2_eff, peers, world};
    const bool f32 = (mode == 2) && cut2_eff >= 1e-3 && cut2_eff <= 100.0;
    if (f32)       nb_scan32_kernel<<<grid, kScanThreads, 0, st>>>(A);
    else if (gram || mode == 2) nb_scan_kernel<1><<<grid, kScanThreads, 0, st>>>(A);
    else           nb_scan_kernel<0><<<grid, kScanThreads, 0, st>>>(A);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Exact force pass (edmd.cpp:258-276, worker.cpp:166-179, master.cpp:297-322).
//
// Work item = (marked tetrad, slot group): ONE WARP owns slot group w of a tetrad that takes part in a flagged
// pair — the spatially compact group of <= 32 atoms the scan's block masks refer to — lane l the atom in slot
// 32w + l.  Items are handed out through an atomic counter, so the warps of the whole GPU stay busy however
// unevenly the contacts are spread (most groups of a marked tetrad touch nothing and cost one adjacency walk).
//   * 32 adjacency entries at a time, each lane fetches one entry's block mask and reduces it to the 8-bit set
//     of PARTNER groups that can reach this warp's group; entries with an empty set are skipped by a ballot;
//   * for a kept partner the warp gathers the atoms of those partner groups in ASCENDING ATOM ORDER (per-set
//     table atom -> group) into its own shared-memory strip, then every lane runs the reference's expression
//     tree against each gathered atom.
// Bit-exactness: a lane adds its atom's contributions in ascending partner-atom order starting from +0.0 per
// pair, then adds the pair partial to the tetrad total in pair-list order — the reference's order.  Skipped
// atom pairs lie outside the effective cutoff and would add +-0.0, which never changes a sum that started at
// +0.0.  d = x(t1) - x(t2) and the sign of the update depend on which end of the pair this tetrad is, but
// F1_i -= d * pf  and  F2_j += d * pf  are both, exactly, "mine += (x_partner - x_mine) * pf" (negation commutes
// with IEEE rounding), so one form serves.  Each atom's force is produced by exactly one lane in a fixed
// order: no atomics on forces, bit-stable, and bit-identical to the W = 1 reference.
constexpr int kAccBuf = 128;    // gathered partner atoms per round (4 chunks of 32)
constexpr int kForceWarps = 4;  // warps per CTA of the force kernel (independent of each other)

struct AccArgs {
    const double* crd;         // [n][3][S]
    const int* natoms;         // [n]
    const int* param_id;       // [n]
    const double* chg;         // [P][S] charges (abq[3i+2])
    const unsigned long long* hit;   // [np] block masks from the distance pass; 0 = pair not flagged
    const long long* adj_off;  // [n+1] CSR over tetrads
    const int* adj_partner;    // partner index; bit 31 set when THIS tetrad is t2 of the pair
    const int* adj_pair;       // pair index (for the mask)
    const int* perm;           // [P][kMaxStride] slot -> atom (-1 = empty slot)
    const unsigned long long* grp;   // [P][32]: byte c of entry l = slot group of atom 32c + l (0xff = no atom)
    double* fnb;               // [n][3][S] out, clipped
    double* epart;             // [n][8][2] out: per slot group, NB energy and electrostatic energy (unclipped)
    const int* list;           // compact list of tetrads to process, or nullptr: tetrads t0 .. t0+count-1
    const unsigned* count;     // device-side length of `list` (nullptr: `count_host`)
    unsigned* ticket;          // work-item counter (zeroed before the launch)
    int t0, count_host;
    int S;
    double cut2_eff;           // see file header
    double krep, qfac;
    int clip;                  // 1: clamp force components to [-1,1] (master.cpp:304-305)
};

struct ForceSmem {
    double px[kForceWarps][kAccBuf], py[kForceWarps][kAccBuf], pz[kForceWarps][kAccBuf], pq[kForceWarps][kAccBuf];
};

template <bool HAS_Q>
__device__ __forceinline__ void acc_partner(const AccArgs& A, ForceSmem& sm, int w, int lane, int u, unsigned pm,
                                            double xi, double yi, double zi, double qi,
                                            double& F0, double& F1, double& F2, double& E0, double& E1) {
    const unsigned full = 0xffffffffu;
    const int S = A.S;
    const int psu = A.param_id[u];
    const double* U = A.crd + (size_t)u * 3 * S;
    const unsigned long long gw = A.grp[(size_t)psu * 32 + lane];
    const double k14 = __dmul_rn(0.25, A.krep), km2 = __dmul_rn(-2.0, A.krep);
    const double q12 = __dmul_rn(0.5, A.qfac), q2 = __dmul_rn(2.0, A.qfac);
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;   // per-pair partial, starts at 0 (edmd.cpp:245-246)
    double pe0 = 0.0, pe1 = 0.0;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        int cnt = 0;
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const int c = 4 * half + c4;
            const unsigned gj = (unsigned)(gw >> (8 * c)) & 0xffu;
            const bool want = gj < 8u && ((pm >> gj) & 1u);
            const unsigned bits = __ballot_sync(full, want);
            if (want) {
                const int j = 32 * c + lane;
                const int pos = cnt + __popc(bits & ((1u << lane) - 1u));
                sm.px[w][pos] = U[j]; sm.py[w][pos] = U[S + j]; sm.pz[w][pos] = U[2 * S + j];
                if (HAS_Q) sm.pq[w][pos] = A.chg[(size_t)psu * S + j];
            }
            cnt += __popc(bits);
        }
        __syncwarp();
#pragma unroll 2
        for (int jj = 0; jj < cnt; jj++) {
            const double dx = __dsub_rn(sm.px[w][jj], xi), dy = __dsub_rn(sm.py[w][jj], yi), dz = __dsub_rn(sm.pz[w][jj], zi);
            double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));   // edmd.cpp:256
            if (r2 < 1e-9) r2 = 1e-9;
            if (r2 < A.cut2_eff) {
                const double d = __dsub_rn(2.0, r2);
                const double a = 0.0 < d ? d : 0.0;                                         // :260
                pe0 = __dadd_rn(pe0, __dmul_rn(__dmul_rn(k14, a), a));                       // :264
                double pf = __dmul_rn(km2, a);                                               // :268
                if (HAS_Q) {
                    const double qq = __dmul_rn(qi, sm.pq[w][jj]);                           // :261 (a product commutes)
                    pe1 = __dadd_rn(pe1, __dmul_rn(__dmul_rn(q12, qq), r2));                 // :265
                    pf = __dsub_rn(pf, __ddiv_rn(__dmul_rn(q2, qq), __dmul_rn(r2, r2)));
                }
                g0 = __dadd_rn(g0, __dmul_rn(dx, pf));                                       // :269-275
                g1 = __dadd_rn(g1, __dmul_rn(dy, pf));
                g2 = __dadd_rn(g2, __dmul_rn(dz, pf));
            }
        }
        __syncwarp();
    }
    F0 = __dadd_rn(F0, g0); F1 = __dadd_rn(F1, g1); F2 = __dadd_rn(F2, g2);   // worker.cpp:173-178
    E0 += pe0; E1 += pe1;
}

template <bool HAS_Q>
__global__ void __launch_bounds__(32 * kForceWarps, 5) nb_force_kernel(AccArgs A) {
    __shared__ ForceSmem sm;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wslot = threadIdx.x >> 5;
    const int S = A.S;
    // (warp-uniform values pass through redux.sync, whose result lives in a uniform register: the loops below
    // then branch on provably uniform values and the ballots and shuffles in them need no reconvergence code)
    const unsigned items = 8u * __reduce_max_sync(full, A.count ? *A.count : (unsigned)A.count_host);
    for (;;) {
        const unsigned item = __reduce_max_sync(full, lane == 0 ? atomicAdd(A.ticket, 1u) : 0u);
        if (item >= items) return;
        const int t = __reduce_max_sync(full, A.list ? A.list[item >> 3] : A.t0 + (int)(item >> 3));
        const int w = (int)(item & 7u);                          // my slot group
        const int ps = A.param_id[t];
        const int a = A.perm[(size_t)ps * kMaxStride + 32 * w + lane];   // my atom
        const bool on = a >= 0;
        const double* X = A.crd + (size_t)t * 3 * S;
        // an empty slot sits far away from everything: it never passes the cutoff test
        double xi = 1e150, yi = 1e150, zi = 1e150, qi = 0.0;
        if (on) {
            xi = X[a]; yi = X[S + a]; zi = X[2 * S + a];
            if (HAS_Q) qi = A.chg[(size_t)ps * S + a];
        }
        double F0 = 0.0, F1 = 0.0, F2 = 0.0;   // tetrad totals, pair-list order
        double E0 = 0.0, E1 = 0.0;             // this lane's share of the pair energies

        // (adjacency offsets fit 32 bits: the builder limits the list to 2^30 pairs)
        const int e_begin = __reduce_max_sync(full, (int)A.adj_off[t]), e_end = __reduce_max_sync(full, (int)A.adj_off[t + 1]);
        for (int base = e_begin; base < e_end; base += 32) {
            const int e = base + lane;
            unsigned pmv = 0; int code = 0;
            if (e < e_end) {
                const unsigned long long m = A.hit[A.adj_pair[e]];
                if (m) {
                    code = A.adj_partner[e];
                    const int u = code & 0x7fffffff;
                    // mask bit 8g + r: group r of the lower-numbered tetrad x group g of the higher-numbered one
                    pmv = t < u ? column_bits(m, w) : (unsigned)(m >> (8 * w)) & 0xffu;
                }
            }
            unsigned todo = __ballot_sync(full, pmv != 0u);
            while (todo) {
                const int b = __ffs(todo) - 1;
                todo &= todo - 1;
                const int u = __reduce_max_sync(full, lane == b ? (code & 0x7fffffff) : 0);
                const unsigned pm = __reduce_max_sync(full, lane == b ? pmv : 0u);
                acc_partner<HAS_Q>(A, sm, wslot, lane, u, pm, xi, yi, zi, qi, F0, F1, F2, E0, E1);
            }
        }

        // clip (master.cpp:304-305) and store
        if (A.clip) {
            F0 = F0 < -1.0 ? -1.0 : (F0 > 1.0 ? 1.0 : F0);
            F1 = F1 < -1.0 ? -1.0 : (F1 > 1.0 ? 1.0 : F1);
            F2 = F2 < -1.0 ? -1.0 : (F2 > 1.0 ? 1.0 : F2);
        }
        double* F = A.fnb + (size_t)t * 3 * S;
        if (on) { F[a] = F0; F[S + a] = F1; F[2 * S + a] = F2; }
        // this group's share of the tetrad's pair energies: fixed-order butterfly, summed over the groups later
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { E0 += __shfl_xor_sync(full, E0, o); E1 += __shfl_xor_sync(full, E1, o); }
        if (lane == 0) { A.epart[((size_t)t * 8 + w) * 2] = E0; A.epart[((size_t)t * 8 + w) * 2 + 1] = E1; }
    }
}

// e_nb[t] = sum over the eight slot groups, in group order (one thread per tetrad)
__global__ void nb_energy_kernel(const double* epart, double* e_nb, int t0, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const size_t t = (size_t)(t0 + i);
    double e0 = 0.0, e1 = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) { e0 += epart[(t * 8 + w) * 2]; e1 += epart[(t * 8 + w) * 2 + 1]; }
    e_nb[2 * t] = e0; e_nb[2 * t + 1] = e1;
}

static void launch_force(const AccArgs& A, int grid, cudaStream_t st) {
    if (A.qfac != 0.0) nb_force_kernel<true><<<grid, 32 * kForceWarps, 0, st>>>(A);
    else               nb_force_kernel<false><<<grid, 32 * kForceWarps, 0, st>>>(A);
}

// Unfused form (edmd_nb_accumulate / edmd_nb_forces): every owned tetrad is a work item.
cudaError_t launch_nb_accumulate(const double* crd, const int* natoms, const int* param_id, const double* chg,
                                 const unsigned long long* hit, const long long* adj_off, const int* adj_partner,
                                 const int* adj_pair, const int* perm, const unsigned long long* grp, double* fnb,
                                 double* e_nb, double* epart, unsigned* ticket, int t0, int count, int S, int sms,
                                 const SimParams& sp, int clip, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(ticket, 0, sizeof(unsigned), st);
    if (err != cudaSuccess) return err;
    AccArgs A{crd, natoms, param_id, chg, hit, adj_off, adj_partner, adj_pair, perm, grp, fnb, epart, nullptr, nullptr,
              ticket, t0, count, S, nb_effective_cut2(sp), sp.krep, sp.qfac, clip};
    const long long warps = 8LL * count;
    const int grid = (int)std::min<long long>((warps + kForceWarps - 1) / kForceWarps, 5LL * sms);
    launch_force(A, grid, st);
    nb_energy_kernel<<<(count + 127) / 128, 128, 0, st>>>(epart, e_nb, t0, count);
    return cudaGetLastError();
}

// ---- the fused tail of the step ---------------------------------------------------------------------
// NB accumulate + clip (master.cpp:297-322), then update_Velocities and update_Coordinates
// (edmd.cpp:289-327).  Other CTAs read a tetrad's coordinates as a partner while it is being integrated, so
// the new coordinates go to a SECOND buffer (I.crd_out); the next ED pass reads them from there.
//
// Almost every tetrad has no flagged pair, and then the tail is a pure streaming update.
//   mark_tetrads_kernel   flagged pair -> both ends marked; the first marker of an owned tetrad appends it to
//                         a compact list (the order of the list varies from run to run, each tetrad's result
//                         does not depend on it);
//   nb_force_kernel       the exact force pass over (marked tetrad, slot group) items, see above;
//   finish_unmarked_kernel  one CTA per owned tetrad without a flagged pair: lean streaming velocity +
//                         coordinate update (zero NB force), 6 CTAs per SM to keep HBM busy;
//   finish_marked_kernel  the same update for the listed tetrads, with the clipped NB force and the pair
//                         energies of the force pass read back.
__global__ void mark_tetrads_kernel(const unsigned long long* hit, const int2* pairs, long long np, int* mark,
                                    int* list, unsigned* count, int t0, int t1) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
        if (!hit[p]) continue;
        const int2 pr = pairs[p];
        if (pr.x >= t0 && pr.x < t1 && atomicExch(&mark[pr.x], 1) == 0) list[atomicAdd(count, 1u)] = pr.x;
        if (pr.y >= t0 && pr.y < t1 && atomicExch(&mark[pr.y], 1) == 0) list[atomicAdd(count, 1u)] = pr.y;
    }
}

__global__ void __launch_bounds__(kThreads, 6) finish_unmarked_kernel(IntArgs I, const int* mark, unsigned char* dirty,
                                                                    double* fnb, double* e_nb, long* bump, long bump_by) {
    __shared__ double red[1][kWarps];
    const int t = I.t0 + blockIdx.x;
    // the random-stream call counter of the replayed step graph advances here (its reader, the generator
    // kernel, ran earlier in the step)
    if (bump && blockIdx.x == 0 && threadIdx.x == 0) *bump += bump_by;
    if (mark[t]) return;
    const int S = I.ps.S;
    if (dirty[t]) {   // forces left by an earlier step in which this tetrad had a flagged pair
        for (int k = threadIdx.x; k < 3 * S; k += blockDim.x) fnb[(size_t)t * 3 * S + k] = 0.0;
        if (threadIdx.x == 0) { e_nb[2 * (size_t)t] = 0.0; e_nb[2 * (size_t)t + 1] = 0.0; dirty[t] = 0; }
    }
    const double zero[3] = {0.0, 0.0, 0.0};
    integrate_tetrad(I, t, zero, red);
}

__global__ void __launch_bounds__(kThreads, 4) finish_marked_kernel(IntArgs I, const int* list, const unsigned* count,
                                                                  unsigned char* dirty, const double* fnb,
                                                                  const double* epart, double* e_nb) {
    __shared__ double red[1][kWarps];
    const int S = I.ps.S;
    const int tid = threadIdx.x;
    const unsigned total = *count;
    for (unsigned i = blockIdx.x; i < total; i += gridDim.x) {
        const int t = list[i];
        double f[3] = {0.0, 0.0, 0.0};
        if (tid < I.ps.na[I.param_id[t]]) {
#pragma unroll
            for (int c = 0; c < 3; c++) f[c] = fnb[(size_t)t * 3 * S + c * S + tid];
        }
        if (tid == 0) {
            double e0 = 0.0, e1 = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) { e0 += epart[((size_t)t * 8 + w) * 2]; e1 += epart[((size_t)t * 8 + w) * 2 + 1]; }
            e_nb[2 * (size_t)t] = e0; e_nb[2 * (size_t)t + 1] = e1;
            dirty[t] = 1;
        }
        integrate_tetrad(I, t, f, red);
    }
}

cudaError_t launch_nb_finish_split(const double* crd_ro, const int* natoms, const int* param_id, const double* chg,
                                   const unsigned long long* hit, const int2* pairs, long long np, const long long* adj_off,
                                   const int* adj_partner, const int* adj_pair, const int* perm,
                                   const unsigned long long* grp, double* fnb, double* e_nb, double* epart,
                                   const ParamSets& ps, double* vel, double* crd_out, const double* fed,
                                   const double* rnd, double* temp, int* mark, unsigned char* dirty, int n,
                                   int t0, int count, int S, int sms, const SimParams& sp, long* bump, long bump_by,
                                   cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    // mark[0..n) flags, then the list counter, the work-item counter and the distance pass's item counter (all
    // zeroed together before the distance pass, engine.cu: do_scan), then the compact list of marked tetrads
    unsigned* cnt = reinterpret_cast<unsigned*>(mark + n);
    unsigned* ticket = reinterpret_cast<unsigned*>(mark + n + 1);
    int* list = mark + n + 4;
    if (np > 0) {
        long long blocks = (np + 255) / 256;
        if (blocks > sms * 16) blocks = sms * 16;
        mark_tetrads_kernel<<<(int)blocks, 256, 0, st>>>(hit, pairs, np, mark, list, cnt, t0, t0 + count);
        // the number of marked tetrads is only known on the device: a grid that fills the SMs draws work items
        AccArgs A{crd_ro, natoms, param_id, chg, hit, adj_off, adj_partner, adj_pair, perm, grp, fnb, epart, list, cnt,
                  ticket, t0, count, S, nb_effective_cut2(sp), sp.krep, sp.qfac, 1};
        const long long warps = 8LL * count;
        const int grid = (int)std::min<long long>((warps + kForceWarps - 1) / kForceWarps, 5LL * sms);
        launch_force(A, grid, st);
    }
    IntArgs I{ps, param_id, vel, crd_ro, crd_out, fed, rnd, nullptr, temp, t0, 1, 1, sp.dt, sp.gamma, sp.tautp, sp.scaled};
    finish_unmarked_kernel<<<count, kThreads, 0, st>>>(I, mark, dirty, fnb, e_nb, bump, bump_by);
    if (np > 0) {
        const int grid = count < 4 * sms ? count : 4 * sms;
        finish_marked_kernel<<<grid, kThreads, 0, st>>>(I, list, cnt, dirty, fnb, epart, e_nb);
    }
    return cudaGetLastError();
}

}  // namespace edmd
