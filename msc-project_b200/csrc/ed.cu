// ed.cu — ED restoring force: QCP superposition + projection on the PCA eigenvectors.
//
// Replaces EDMD::calculate_ED_Forces (/root/reference/src/edmd.cpp:64-166) and the QCP
// routines it calls (src/qcprot/qcprot.c:90-407), batched over tetrads.
//
// Two kernels per batch:
//
//  superpose_kernel (one warp per tetrad, ~20 tetrads resident per SM): centroid, 3x3
//    cross-covariance, QCP rotation and the offset vector.  The rotation feeds
//    d = R*x_c - avg_c, a difference of ~10 A vectors that is ~0.1 A long, so last-bit
//    differences of the sums are amplified ~100x into the forces.  To stay robustly inside the
//    1e-10 parity bound (and usually at 1e-15) the reductions are done in the reference's
//    SEQUENTIAL atom order (qcprot.c:370-379, 132-157; edmd.cpp:97-101): per-atom terms are
//    produced 32 atoms at a time by the whole warp into a shared tile, then lane k runs the
//    k-th running sum as a chain of DADDs.  The sums that depend only on the parameter set
//    (centroid and norm of the average structure) are computed once at upload.  This file is
//    compiled with -fmad=false, so every expression rounds exactly like the CPU build and R,
//    centroids and offset are BIT-IDENTICAL to the reference's.  Latency-bound (3 x 252
//    dependent adds per tetrad) and hidden by the independent warps of an SM.  (The offset vector as a
//    tree sum — it only translates the tetrad — was tried: 0.10 ms faster at 1e5 tetrads, but the 1-ulp
//    shift of a tetrad's coordinates (1e-14 A instead of 1e-17) is amplified by the cancellation in
//    clashing electrostatic contacts to 4e-10 on NB forces: over the parity bound, dropped.)
//
//  (Large systems whose tetrads share parameter sets with <= 12 eigenvectors use ed_project_mma_kernel, ed_mma.cu: the
//  same two passes as FP64 tensor-core contractions over groups of eight tetrads.  The two kernels below serve small
//  systems, sets with more eigenvectors and edmd_set_ed_kernel(1|2).)
//  ed_project_warp_kernel (parameter sets with <= 12 eigenvectors: one warp per tetrad, the set's
//    eigenvectors staged in shared memory once per run of tetrads) and ed_project_kernel (any
//    number of eigenvectors: 256 threads per tetrad, thread a <-> atom a, coefficients in
//    registers): displacement, projections on the eigenvectors (fixed-order tree sums — the only
//    place the summation order differs from the CPU, ~1e-16 relative), re-embedding, force,
//    rotation back.  The 72.6 KB eigenvector block crosses the memory system once per run (the
//    reference walks it twice per tetrad, the second time with stride 3N, edmd.cpp:106-127).
//    Roofline: HBM — 96,872 algorithmic bytes per tetrad (SURVEY.md §8a a-ED); with
//    de-duplicated parameter sets resident in L2 / shared memory the DRAM traffic drops to the
//    24 KB of state.
#include "common.cuh"
#include "kernels.h"

namespace edmd {

// ---- QCP rotation ---------------------------------------------------------------------------------
// qcp_rotation below restates, expression for expression (bit-identical rotation matrices are the
// contract), the closed-form algebra of FastCalcRMSDAndRotation in the QCP reference implementation
// qcprot.c v1.4, which carries this notice:
//
//   Copyright (c) 2009-2013 Pu Liu and Douglas L. Theobald.  All rights reserved.
//   Redistribution and use in source and binary forms, with or without modification, are permitted
//   provided that the following conditions are met:
//   * Redistributions of source code must retain the above copyright notice, this list of conditions
//     and the following disclaimer.
//   * Redistributions in binary form must reproduce the above copyright notice, this list of
//     conditions and the following disclaimer in the documentation and/or other materials provided
//     with the distribution.
//   * Neither the name of the <ORGANIZATION> nor the names of its contributors may be used to endorse
//     or promote products derived from this software without specific prior written permission.
//   THIS SOFTWARE IS PROVIDED BY THE COPYRIGHT HOLDERS AND CONTRIBUTORS "AS IS" AND ANY EXPRESS OR
//   IMPLIED WARRANTIES, INCLUDING, BUT NOT LIMITED TO, THE IMPLIED WARRANTIES OF MERCHANTABILITY AND
//   FITNESS FOR A PARTICULAR PURPOSE ARE DISCLAIMED.  IN NO EVENT SHALL THE COPYRIGHT HOLDER OR
//   CONTRIBUTORS BE LIABLE FOR ANY DIRECT, INDIRECT, INCIDENTAL, SPECIAL, EXEMPLARY, OR CONSEQUENTIAL
//   DAMAGES (INCLUDING, BUT NOT LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS OR SERVICES; LOSS OF USE,
//   DATA, OR PROFITS; OR BUSINESS INTERRUPTION) HOWEVER CAUSED AND ON ANY THEORY OF LIABILITY, WHETHER
//   IN CONTRACT, STRICT LIABILITY, OR TORT (INCLUDING NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY OUT OF
//   THE USE OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF SUCH DAMAGE.
//
//   Method: D. L. Theobald (2005) Acta Cryst. A 61(4):478-480; P. Liu, D. K. Agrafiotis and
//   D. L. Theobald (2009) J. Comput. Chem. 31(7):1561-1563.
//
// Rotation from the 3x3 cross-covariance M (row-major, ref (x) mov) and e0 = (G_ref+G_mov)/2:
// largest root of the QCP quartic by Newton from e0 (stop |dl| < 1e-11 |l|, <= 50 iterations),
// quaternion = a non-vanishing column of adj(K - l I) (norm^2 >= 1e-6, else next column, else
// identity) — same decisions as FastCalcRMSDAndRotation, qcprot.c:164-340.
__device__ void qcp_rotation(double R[9], const double M[9], double e0) {
    const double xx = M[0], xy = M[1], xz = M[2];
    const double yx = M[3], yy = M[4], yz = M[5];
    const double zx = M[6], zy = M[7], zz = M[8];
    const double xx2 = xx * xx, yy2 = yy * yy, zz2 = zz * zz;
    const double xy2 = xy * xy, yz2 = yz * yz, xz2 = xz * xz;
    const double yx2 = yx * yx, zy2 = zy * zy, zx2 = zx * zx;

    const double ta = 2.0 * (yz * zy - yy * zz);
    const double tb = yy2 + zz2 - xx2 + yz2 + zy2;
    const double c2 = -2.0 * (xx2 + yy2 + zz2 + xy2 + yx2 + xz2 + zx2 + yz2 + zy2);
    const double c1 = 8.0 * (xx * yz * zy + yy * zx * xz + zz * xy * yx - xx * yy * zz - yz * zx * xy - zy * yx * xz);
    const double xzp = xz + zx, yzp = yz + zy, xyp = xy + yx;
    const double yzm = yz - zy, xzm = xz - zx, xym = xy - yx;
    const double dp = xx + yy, dm = xx - yy;
    const double tc = xy2 + xz2 - yx2 - zx2;
    const double c0 = tc * tc + (tb + ta) * (tb - ta)
        + (-(xzp) * (yzm) + (xym) * (dm - zz)) * (-(xzm) * (yzp) + (xym) * (dm + zz))
        + (-(xzp) * (yzp) - (xyp) * (dp - zz)) * (-(xzm) * (yzm) - (xyp) * (dp + zz))
        + (+(xyp) * (yzp) + (xzp) * (dm + zz)) * (-(xym) * (yzm) + (xzp) * (dp + zz))
        + (+(xyp) * (yzm) + (xzm) * (dm - zz)) * (-(xym) * (yzp) + (xzm) * (dp - zz));

    double lam = e0;
    for (int it = 0; it < 50; it++) {
        const double prev = lam;
        const double l2 = lam * lam;
        const double b = (l2 + c2) * lam;
        const double a = b + c1;
        lam -= (a * lam + c0) / (2.0 * l2 * lam + b + a);
        if (fabs(lam - prev) < fabs(1e-11 * lam)) break;
    }

    const double a11 = dp + zz - lam, a12 = yzm, a13 = -xzm, a14 = xym;
    const double a21 = yzm, a22 = dm - zz - lam, a23 = xyp, a24 = xzp;
    const double a31 = a13, a32 = a23, a33 = yy - xx - zz - lam, a34 = yzp;
    const double a41 = a14, a42 = a24, a43 = a34, a44 = zz - dp - lam;
    const double m3344 = a33 * a44 - a43 * a34, m3244 = a32 * a44 - a42 * a34;
    const double m3243 = a32 * a43 - a42 * a33, m3143 = a31 * a43 - a41 * a33;
    const double m3144 = a31 * a44 - a41 * a34, m3142 = a31 * a42 - a41 * a32;
    double q1 =  a22 * m3344 - a23 * m3244 + a24 * m3243;
    double q2 = -a21 * m3344 + a23 * m3144 - a24 * m3143;
    double q3 =  a21 * m3244 - a22 * m3144 + a24 * m3142;
    double q4 = -a21 * m3243 + a22 * m3143 - a23 * m3142;
    double qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
    if (qq < 1e-6) {
        q1 =  a12 * m3344 - a13 * m3244 + a14 * m3243;
        q2 = -a11 * m3344 + a13 * m3144 - a14 * m3143;
        q3 =  a11 * m3244 - a12 * m3144 + a14 * m3142;
        q4 = -a11 * m3243 + a12 * m3143 - a13 * m3142;
        qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
        if (qq < 1e-6) {
            const double m1324 = a13 * a24 - a14 * a23, m1224 = a12 * a24 - a14 * a22;
            const double m1223 = a12 * a23 - a13 * a22, m1124 = a11 * a24 - a14 * a21;
            const double m1123 = a11 * a23 - a13 * a21, m1122 = a11 * a22 - a12 * a21;
            q1 =  a42 * m1324 - a43 * m1224 + a44 * m1223;
            q2 = -a41 * m1324 + a43 * m1124 - a44 * m1123;
            q3 =  a41 * m1224 - a42 * m1124 + a44 * m1122;
            q4 = -a41 * m1223 + a42 * m1123 - a43 * m1122;
            qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
            if (qq < 1e-6) {
                q1 =  a32 * m1324 - a33 * m1224 + a34 * m1223;
                q2 = -a31 * m1324 + a33 * m1124 - a34 * m1123;
                q3 =  a31 * m1224 - a32 * m1124 + a34 * m1122;
                q4 = -a31 * m1223 + a32 * m1123 - a33 * m1122;
                qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
                if (qq < 1e-6) {
                    R[0] = R[4] = R[8] = 1.0;
                    R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0.0;
                    return;
                }
            }
        }
    }
    const double nq = sqrt(qq);
    q1 /= nq; q2 /= nq; q3 /= nq; q4 /= nq;
    const double w2 = q1 * q1, i2 = q2 * q2, j2 = q3 * q3, k2 = q4 * q4;
    const double ij = q2 * q3, wk = q1 * q4, ki = q4 * q2, wj = q1 * q3, jk = q3 * q4, wi = q1 * q2;
    R[0] = w2 + i2 - j2 - k2; R[1] = 2 * (ij + wk);     R[2] = 2 * (ki - wj);
    R[3] = 2 * (ij - wk);     R[4] = w2 - i2 + j2 - k2; R[5] = 2 * (jk + wi);
    R[6] = 2 * (ki + wj);     R[7] = 2 * (jk - wi);     R[8] = w2 - i2 - j2 + k2;
}

constexpr int kSupWarps = 4;                      // one warp = one tetrad; no CTA-wide barrier anywhere
constexpr int kSupWords = 24;   // per tetrad: R[9], shift[3], avg centroid[3], x centroid[3], pad
constexpr int kSupRow = kMaxStride + 1;
constexpr int kTileRow = 33;                      // 32 atoms + 1: chain lanes k = 0..9 hit distinct banks

struct SupArgs {
    ParamSets ps;
    const int* param_id;
    const double* crd;   // [n][3][S]
    double* sup;         // [n][kSupWords] out
    int t0, count;
};

// Running sum of p[0..na) in atom order, starting from 0.0 — the reference's `s += p[i]` loops.
// Eight loads are issued before their eight dependent adds, so the chain runs at DADD latency.
__device__ __forceinline__ double seq_sum1(const double* p, int na) {
    double s = 0.0;
    int i = 0;
    for (; i + 8 <= na; i += 8) {
        const double v0 = p[i], v1 = p[i + 1], v2 = p[i + 2], v3 = p[i + 3];
        const double v4 = p[i + 4], v5 = p[i + 5], v6 = p[i + 6], v7 = p[i + 7];
        s = s + v0; s = s + v1; s = s + v2; s = s + v3; s = s + v4; s = s + v5; s = s + v6; s = s + v7;
    }
    for (; i < na; i++) s = s + p[i];
    return s;
}
// s += row[0..cnt) for one 32-atom tile row (cnt <= 32).
__device__ __forceinline__ double tile_chain(double s, const double* row, int cnt) {
    if (cnt == 32) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            double v[16];
#pragma unroll
            for (int u = 0; u < 16; u++) v[u] = row[16 * h + u];
#pragma unroll
            for (int u = 0; u < 16; u++) s = s + v[u];
        }
    } else {
        for (int i = 0; i < cnt; i++) s = s + row[i];
    }
    return s;
}

// Per parameter set, once at upload: centroid of the average structure and G_ref = sum |avg_c|^2
// (CenterCoords + the G1 sum of InnerProduct, qcprot.c:370-386,132-157) — the same for every tetrad
// that shares the set, so superpose_kernel does not repeat these four 252-step chains.
__global__ void ps_precompute_kernel(ParamSets ps, int P, double* pre) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const int S = ps.S, na = ps.na[p];
    const double* AV = ps.avg + (size_t)p * 3 * S;
    const double s = seq_sum1(AV + (size_t)(lane < 3 ? lane : 0) * S, na);
    const double c = s / na;
    const double c0 = __shfl_sync(0xffffffffu, c, 0), c1 = __shfl_sync(0xffffffffu, c, 1), c2 = __shfl_sync(0xffffffffu, c, 2);
    double g = 0.0;
    for (int i = 0; i < na; i++) {
        const double a0 = AV[i] - c0, a1 = AV[S + i] - c1, a2 = AV[2 * S + i] - c2;
        g = g + ((a0 * a0 + a1 * a1) + a2 * a2);
    }
    if (lane < 3) pre[4 * (size_t)p + lane] = c;
    if (lane == 3) pre[4 * (size_t)p + 3] = g;
}
cudaError_t launch_ps_precompute(const ParamSets& ps, int P, double* pre, cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    ps_precompute_kernel<<<(P + 3) / 4, 128, 0, st>>>(ps, P, pre);
    return cudaGetLastError();
}

// One warp per tetrad.  The sequential sums of the reference (centroid, 3x3 covariance + norm,
// offset vector) are 252-step dependent DADD chains; each of them runs in ONE lane, the chains of a
// phase side by side in lanes 0..9, and many warps per SM hide the 8-cycle add latency.  The
// per-atom terms that feed a chain are produced 32 atoms at a time by all 32 lanes and handed over
// through a [10][33] shared tile, so a chain step is one LDS + one DADD.
struct SupSmem {
    double x[3][kSupRow];          // coordinates, centred in place
    double tile[10][kTileRow];
};

__global__ void __launch_bounds__(32 * kSupWarps, 6) superpose_kernel(SupArgs A) {
    __shared__ SupSmem sm_all[kSupWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * kSupWarps + warp;
    if (e >= A.count) return;
    SupSmem& sm = sm_all[warp];
    const unsigned full = 0xffffffffu;
    const int S = A.ps.S;
    const int t = A.t0 + e;
    const int ps = A.param_id[t];
    const int na = A.ps.na[ps];
    const double* X = A.crd + (size_t)t * 3 * S;
    const double* AV = A.ps.avg + (size_t)ps * 3 * S;
    const double* PRE = A.ps.pre + 4 * (size_t)ps;
    const double ca0 = PRE[0], ca1 = PRE[1], ca2 = PRE[2], g_ref = PRE[3];

#pragma unroll
    for (int r = 0; r < kMaxStride / 32; r++) {
        const int i = lane + 32 * r;
        if (i < na) { sm.x[0][i] = X[i]; sm.x[1][i] = X[S + i]; sm.x[2][i] = X[2 * S + i]; }
    }
    __syncwarp();
    // global operands of the next 32-atom block are fetched one block ahead of their use, so their
    // L2 latency hides behind the running chains
    double pa0 = 0.0, pa1 = 0.0, pa2 = 0.0;
    if (lane < na) { pa0 = AV[lane]; pa1 = AV[S + lane]; pa2 = AV[2 * S + lane]; }
    double cx0, cx1, cx2;
    {   // CenterCoords, qcprot.c:370-379: sum in atom order, then divide by len
        const double s = seq_sum1(sm.x[lane < 3 ? lane : 0], na);
        const double c = s / na;
        cx0 = __shfl_sync(full, c, 0); cx1 = __shfl_sync(full, c, 1); cx2 = __shfl_sync(full, c, 2);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kMaxStride / 32; r++) {   // centre in place (qcprot.c:382-386)
        const int i = lane + 32 * r;
        if (i < na) { sm.x[0][i] = sm.x[0][i] - cx0; sm.x[1][i] = sm.x[1][i] - cx1; sm.x[2][i] = sm.x[2][i] - cx2; }
    }
    __syncwarp();

    double R[9];
    double px0 = 0.0, px1 = 0.0, px2 = 0.0;
    {   // InnerProduct, qcprot.c:132-157: chains 0-8 the 3x3 products avg_c (x) x_c, chain 9 the norm of x_c
        double acc = 0.0;
        for (int b = 0; b < na; b += 32) {
            const int i = b + lane;
            double v[10];
#pragma unroll
            for (int k = 0; k < 10; k++) v[k] = 0.0;
            const double r0 = pa0, r1 = pa1, r2 = pa2;
            if (i + 32 < na) { pa0 = AV[i + 32]; pa1 = AV[S + i + 32]; pa2 = AV[2 * S + i + 32]; }
            if (i < na) {
                const double a0 = r0 - ca0, a1 = r1 - ca1, a2 = r2 - ca2;
                const double x0 = sm.x[0][i], x1 = sm.x[1][i], x2 = sm.x[2][i];
                v[0] = a0 * x0; v[1] = a0 * x1; v[2] = a0 * x2;
                v[3] = a1 * x0; v[4] = a1 * x1; v[5] = a1 * x2;
                v[6] = a2 * x0; v[7] = a2 * x1; v[8] = a2 * x2;
                v[9] = (x0 * x0 + x1 * x1) + x2 * x2;
            }
#pragma unroll
            for (int k = 0; k < 10; k++) sm.tile[k][lane] = v[k];
            __syncwarp();
            const int cnt = na - b < 32 ? na - b : 32;
            if (lane < 10) acc = tile_chain(acc, sm.tile[lane], cnt);
            __syncwarp();
        }
        double m[9];
#pragma unroll
        for (int j = 0; j < 9; j++) m[j] = __shfl_sync(full, acc, j);
        const double g_x = __shfl_sync(full, acc, 9);
        if (lane < na) {   // first block of the offset pass, in flight during the quartic solve
            pa0 = AV[lane]; pa1 = AV[S + lane]; pa2 = AV[2 * S + lane];
            px0 = X[lane]; px1 = X[S + lane]; px2 = X[2 * S + lane];
        }
        qcp_rotation(R, m, (g_ref + g_x) * 0.5);
    }

    double off = 0.0;   // offset-vector terms on the UNCENTRED arrays (re-read, L2-resident), edmd.cpp:97-101
    for (int b = 0; b < na; b += 32) {
        const int i = b + lane;
        double w0 = 0.0, w1 = 0.0, w2 = 0.0;
        const double a0 = pa0, a1 = pa1, a2 = pa2, x0 = px0, x1 = px1, x2 = px2;
        if (i + 32 < na) {
            pa0 = AV[i + 32]; pa1 = AV[S + i + 32]; pa2 = AV[2 * S + i + 32];
            px0 = X[i + 32]; px1 = X[S + i + 32]; px2 = X[2 * S + i + 32];
        }
        if (i < na) {
            w0 = (a0 - (R[0] * x0 + R[1] * x1 + R[2] * x2));
            w1 = (a1 - (R[3] * x0 + R[4] * x1 + R[5] * x2));
            w2 = (a2 - (R[6] * x0 + R[7] * x1 + R[8] * x2));
        }
        sm.tile[0][lane] = w0; sm.tile[1][lane] = w1; sm.tile[2][lane] = w2;
        __syncwarp();
        const int cnt = na - b < 32 ? na - b : 32;
        if (lane < 3) off = tile_chain(off, sm.tile[lane], cnt);
        __syncwarp();
    }

    double* out = A.sup + (size_t)t * kSupWords;
#pragma unroll
    for (int j = 0; j < 9; j++) if (lane == j) out[j] = R[j];
    if (lane < 3) {
        out[9 + lane] = off / na;                                            // edmd.cpp:102
        out[12 + lane] = lane == 0 ? ca0 : lane == 1 ? ca1 : ca2;
        out[15 + lane] = lane == 0 ? cx0 : lane == 1 ? cx1 : cx2;
    }
}

// ---- two tetrads per warp ----------------------------------------------------------------------------
// Same sums in the same order, but a warp carries TWO tetrads, one per half-warp: the per-atom terms are
// produced 16 atoms at a time by each half into its own [10][17] shared tile, and the chains of both tetrads
// advance together (lanes 0..9 and 16..25), so one chain instruction serves two tetrads.  Nothing else is
// staged in shared memory — the coordinates are re-read from L1/L2 in each of the three passes, one block
// ahead of their use — so a CTA needs 11 KB and the SM holds 64 tetrads in flight instead of 24: the chains are
// bound by the latency of dependent adds, and twice the independent chains per issued instruction plus 2.7x
// the tetrads in flight is what this layout buys.
constexpr int kS2Warps = 4;
constexpr int kS2Row = 17;
struct Sup2Tile { double v[2][10][kS2Row]; };

// s += row[0..cnt) (cnt <= 16), sequentially
__device__ __forceinline__ double tile_chain16(double s, const double* row, int cnt) {
    if (cnt == 16) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; u++) v[u] = row[u];
#pragma unroll
        for (int u = 0; u < 16; u++) s = s + v[u];
    } else {
        for (int i = 0; i < cnt; i++) s = s + row[i];
    }
    return s;
}

// Instruction count is what bounds this kernel (ncu: 67 % of the issue slots, FP64 pipe 30 %), so the loops are
// written for it: the atom stride is a template constant when it is the usual 256 (the three planes of a block
// are then one address plus immediates), the prefetch index is clamped instead of guarded, and tile entries of
// atoms beyond the tetrad's count are left as they come — the chains never read them.
template <int SC>
__global__ void __launch_bounds__(32 * kS2Warps, 8) superpose2_kernel(SupArgs A) {
    __shared__ Sup2Tile tiles[kS2Warps];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, hl = lane & 15, base = lane & 16;
    double (*tile)[kS2Row] = tiles[warp].v[half];
    const int S = SC ? SC : A.ps.S;
    // warps walk the tetrad pairs with a grid stride; the launcher sizes the grid so that every warp gets the same
    // number of pairs (+-1) instead of leaving a thin second wave on a few SMs
    const int units = (A.count + 1) >> 1, stride = gridDim.x * kS2Warps;
    for (int unit = __reduce_max_sync(full, blockIdx.x * kS2Warps + warp); unit < units; unit += stride) {
    const int e = unit * 2 + half;
    const bool live = e < A.count;                       // the last pair may hold a single tetrad
    const int t = A.t0 + (live ? e : A.count - 1);       // (a dead half reads a valid tetrad and writes nothing)
    const int ps = A.param_id[t];
    const int na = live ? A.ps.na[ps] : 0;
    // both halves walk the same number of blocks (redux.sync: the result is in a uniform register, so the block
    // loops branch on a provably warp-uniform value and their __syncwarp()s cost nothing extra)
    const int na_max = __reduce_max_sync(full, na);
    const double* X = A.crd + (size_t)t * 3 * S;
    const double* AV = A.ps.avg + (size_t)ps * 3 * S;
    const double* PRE = A.ps.pre + 4 * (size_t)ps;
    const double ca0 = PRE[0], ca1 = PRE[1], ca2 = PRE[2], g_ref = PRE[3];
    const int last = S - 1;                              // prefetches past the tetrad read its zero padding

    // pass 1 — CenterCoords, qcprot.c:370-379: sums in atom order, then divide by len
    double cx0, cx1, cx2;
    {
        double acc = 0.0;
        double p0 = X[hl], p1 = X[S + hl], p2 = X[2 * S + hl];
        for (int b = 0; b < na_max; b += 16) {
            tile[0][hl] = p0; tile[1][hl] = p1; tile[2][hl] = p2;
            const int nx = min(b + 16 + hl, last);
            p0 = X[nx]; p1 = X[S + nx]; p2 = X[2 * S + nx];
            __syncwarp();
            const int cnt = na - b;
            if (hl < 3 && cnt > 0) acc = tile_chain16(acc, tile[hl], cnt < 16 ? cnt : 16);
            __syncwarp();
        }
        const double c = na > 0 ? acc / na : 0.0;
        cx0 = __shfl_sync(full, c, base); cx1 = __shfl_sync(full, c, base + 1); cx2 = __shfl_sync(full, c, base + 2);
    }

    // pass 2 — InnerProduct, qcprot.c:132-157: chains 0-8 the 3x3 products avg_c (x) x_c, chain 9 the norm of x_c
    double R[9];
    {
        double acc = 0.0;
        double pa0 = AV[hl], pa1 = AV[S + hl], pa2 = AV[2 * S + hl];
        double px0 = X[hl], px1 = X[S + hl], px2 = X[2 * S + hl];
        for (int b = 0; b < na_max; b += 16) {
            const double a0 = pa0 - ca0, a1 = pa1 - ca1, a2 = pa2 - ca2;
            const double x0 = px0 - cx0, x1 = px1 - cx1, x2 = px2 - cx2;      // centred copies, qcprot.c:382-386
            const int nx = min(b + 16 + hl, last);
            pa0 = AV[nx]; pa1 = AV[S + nx]; pa2 = AV[2 * S + nx];
            px0 = X[nx]; px1 = X[S + nx]; px2 = X[2 * S + nx];
            tile[0][hl] = a0 * x0; tile[1][hl] = a0 * x1; tile[2][hl] = a0 * x2;
            tile[3][hl] = a1 * x0; tile[4][hl] = a1 * x1; tile[5][hl] = a1 * x2;
            tile[6][hl] = a2 * x0; tile[7][hl] = a2 * x1; tile[8][hl] = a2 * x2;
            tile[9][hl] = (x0 * x0 + x1 * x1) + x2 * x2;
            __syncwarp();
            const int cnt = na - b;
            if (hl < 10 && cnt > 0) acc = tile_chain16(acc, tile[hl], cnt < 16 ? cnt : 16);
            __syncwarp();
        }
        double m[9];
#pragma unroll
        for (int j = 0; j < 9; j++) m[j] = __shfl_sync(full, acc, base + j);
        const double g_x = __shfl_sync(full, acc, base + 9);
        qcp_rotation(R, m, (g_ref + g_x) * 0.5);
    }

    // pass 3 — offset-vector terms on the UNCENTRED arrays, edmd.cpp:97-101
    double off = 0.0;
    {
        double pa0 = AV[hl], pa1 = AV[S + hl], pa2 = AV[2 * S + hl];
        double px0 = X[hl], px1 = X[S + hl], px2 = X[2 * S + hl];
        for (int b = 0; b < na_max; b += 16) {
            tile[0][hl] = (pa0 - (R[0] * px0 + R[1] * px1 + R[2] * px2));
            tile[1][hl] = (pa1 - (R[3] * px0 + R[4] * px1 + R[5] * px2));
            tile[2][hl] = (pa2 - (R[6] * px0 + R[7] * px1 + R[8] * px2));
            const int nx = min(b + 16 + hl, last);
            pa0 = AV[nx]; pa1 = AV[S + nx]; pa2 = AV[2 * S + nx];
            px0 = X[nx]; px1 = X[S + nx]; px2 = X[2 * S + nx];
            __syncwarp();
            const int cnt = na - b;
            if (hl < 3 && cnt > 0) off = tile_chain16(off, tile[hl], cnt < 16 ? cnt : 16);
            __syncwarp();
        }
    }

    if (live) {
        double* out = A.sup + (size_t)t * kSupWords;
#pragma unroll
        for (int j = 0; j < 9; j++) if (hl == j) out[j] = R[j];
        if (hl < 3) {
            out[9 + hl] = off / na;                                              // edmd.cpp:102
            out[12 + hl] = hl == 0 ? ca0 : hl == 1 ? ca1 : ca2;
            out[15 + hl] = hl == 0 ? cx0 : hl == 1 ? cx1 : cx2;
        }
    }
    }   // tetrad pairs of this warp
}

// Sum 12 per-lane values over the warp with a halving butterfly: at each of the 5 stages a lane
// keeps half of its values and hands the other half to its partner, so the 12 sums cost
// 6+3+2+1+1 = 13 adds and exchanges per lane instead of 12 x 5.  On return lane l holds the warp
// total of value index reduce12_index(l) (or garbage where that is -1).  Fixed order -> bit-stable.
__device__ __forceinline__ int reduce12_index(int lane) {
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    if (b2 && b1) return -1;
    return 6 * b4 + 3 * b3 + (b2 ? 2 : b1);
}
__device__ __forceinline__ double warp_reduce12(const double (&v)[12], int lane) {
    const unsigned full = 0xffffffffu;
    double a[6];
    {   // stage 16: 12 -> 6
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const double keep = hi ? v[i + 6] : v[i], send = hi ? v[i] : v[i + 6];
            a[i] = keep + __shfl_xor_sync(full, send, 16);
        }
    }
    double b[3];
    {   // stage 8: 6 -> 3
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double keep = hi ? a[i + 3] : a[i], send = hi ? a[i] : a[i + 3];
            b[i] = keep + __shfl_xor_sync(full, send, 8);
        }
    }
    double c[2];
    {   // stage 4: 3 -> 2 (value 3 is padding)
        const bool hi = lane & 4;
        const double k0 = hi ? b[2] : b[0], s0 = hi ? b[0] : b[2];
        const double k1 = hi ? 0.0 : b[1], s1 = hi ? b[1] : 0.0;
        c[0] = k0 + __shfl_xor_sync(full, s0, 4);
        c[1] = k1 + __shfl_xor_sync(full, s1, 4);
    }
    double d;
    {   // stage 2: 2 -> 1
        const bool hi = lane & 2;
        const double keep = hi ? c[1] : c[0], send = hi ? c[0] : c[1];
        d = keep + __shfl_xor_sync(full, send, 2);
    }
    return d + __shfl_xor_sync(full, d, 1);   // stage 1
}

struct EdArgs {
    ParamSets ps;
    const int* param_id;
    const double* crd_in;  // [n][3][S] current coordinates
    double* crd;         // [n][3][S] out: the re-embedded coordinates (may alias crd_in)
    double* fed;         // [n][3][S] out
    double* e_ed;        // [n] out
    const double* sup;   // [n][kSupWords] from superpose_kernel
    int t0;              // first tetrad of this launch
    double scaled;
    const int* order;    // [count] tetrads of this launch sorted by parameter set
    int count;
    int chunk;           // tetrads per CTA
};

__global__ void __launch_bounds__(kThreads, 2) ed_project_kernel(EdArgs A) {
    __shared__ double red[16][kWarps];
    __shared__ double sProj[64];
    __shared__ double sPw[64];

    const int tid = threadIdx.x;
    const int S = A.ps.S;
    const int e0 = blockIdx.x * A.chunk;
    const int e1 = (e0 + A.chunk < A.count) ? e0 + A.chunk : A.count;

    // A CTA walks a run of tetrads that (mostly) share one parameter set: the thread's 3 x 12
    // eigenvector coefficients and its atom of the average structure stay in registers across the
    // run, so a shared parameter table is read once per run instead of once per tetrad.
    int cur_ps = -1, na = 0, nev = 0;
    bool on = false;
    double a0 = 0, a1 = 0, a2 = 0;
    double ev[kNevReg][3];
    const double* EV = nullptr;
    const double* EL = nullptr;

    for (int e = e0; e < e1; e++) {
        const int t = A.order[e];
        const int ps = A.param_id[t];
        if (ps != cur_ps) {
            cur_ps = ps;
            na = A.ps.na[ps]; nev = A.ps.nev[ps];
            on = tid < na;
            const double* AV = A.ps.avg + (size_t)ps * 3 * S;
            EV = A.ps.evec + (size_t)ps * A.ps.NEV * 3 * S;
            EL = A.ps.eval + (size_t)ps * A.ps.NEV;
            a0 = a1 = a2 = 0.0;
            if (on) { a0 = AV[tid]; a1 = AV[S + tid]; a2 = AV[2 * S + tid]; }
            const int nchunk0 = nev < kNevReg ? nev : kNevReg;
#pragma unroll
            for (int j = 0; j < kNevReg; j++) {
                ev[j][0] = ev[j][1] = ev[j][2] = 0.0;
                if (on && j < nchunk0) {
                    const double* c = EV + (size_t)j * 3 * S;
                    ev[j][0] = c[tid]; ev[j][1] = c[S + tid]; ev[j][2] = c[2 * S + tid];
                }
            }
        } else if (nev > kNevReg) {   // multi-chunk sets end a tetrad with the LAST chunk in registers
#pragma unroll
            for (int j = 0; j < kNevReg; j++) {
                ev[j][0] = ev[j][1] = ev[j][2] = 0.0;
                if (on) {
                    const double* c = EV + (size_t)j * 3 * S;
                    ev[j][0] = c[tid]; ev[j][1] = c[S + tid]; ev[j][2] = c[2 * S + tid];
                }
            }
        }
        double* X = A.crd + (size_t)t * 3 * S;
        const double* XIN = A.crd_in + (size_t)t * 3 * S;
        const double* SUP = A.sup + (size_t)t * kSupWords;

        double x0 = 0, x1 = 0, x2 = 0;
        if (on) { x0 = XIN[tid]; x1 = XIN[S + tid]; x2 = XIN[2 * S + tid]; }
        double R[9];
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = SUP[k];
        const double shift0 = SUP[9], shift1 = SUP[10], shift2 = SUP[11];

        // displacement in the reference frame, edmd.cpp:89-94 (centred copies as qcprot.c:382-386)
        double d0 = 0, d1 = 0, d2 = 0;
        if (on) {
            const double ac0 = a0 - SUP[12], ac1 = a1 - SUP[13], ac2 = a2 - SUP[14];
            const double xc0 = x0 - SUP[15], xc1 = x1 - SUP[16], xc2 = x2 - SUP[17];
            d0 = R[0] * xc0 + R[1] * xc1 + R[2] * xc2 - ac0;
            d1 = R[3] * xc0 + R[4] * xc1 + R[5] * xc2 - ac1;
            d2 = R[6] * xc0 + R[7] * xc1 + R[8] * xc2 - ac2;
        }

        // projections (edmd.cpp:106-110), kNevReg eigenvectors per pass
        for (int j0 = 0; j0 < nev; j0 += kNevReg) {
            if (j0 > 0) {
#pragma unroll
                for (int j = 0; j < kNevReg; j++) {
                    ev[j][0] = ev[j][1] = ev[j][2] = 0.0;
                    if (on && j0 + j < nev) {
                        const double* c = EV + (size_t)(j0 + j) * 3 * S;
                        ev[j][0] = c[tid]; ev[j][1] = c[S + tid]; ev[j][2] = c[2 * S + tid];
                    }
                }
            }
            double v[kNevReg];
#pragma unroll
            for (int j = 0; j < kNevReg; j++) v[j] = fma(ev[j][0], d0, fma(ev[j][1], d1, ev[j][2] * d2));
            {   // CTA sum of the 12 partial projections: warp butterfly, then 12 threads add the 8 warps
                const int lane = tid & 31, warp = tid >> 5;
                const double w = warp_reduce12(v, lane);
                const int idx = reduce12_index(lane);
                if (idx >= 0 && !(lane & 1)) red[idx][warp] = w;
                __syncthreads();
                if (tid < kNevReg && j0 + tid < nev) {
                    double p = red[tid][0];
#pragma unroll
                    for (int q = 1; q < kWarps; q++) p += red[tid][q];
                    sProj[j0 + tid] = p;
                    sPw[j0 + tid] = p * A.scaled / EL[j0 + tid];
                }
                if (j0 + kNevReg < nev) __syncthreads();   // red[] is reused by the next chunk
            }
        }
        __syncthreads();

        // re-embedding and force in the reference frame (edmd.cpp:113-127)
        double s0 = a0, s1 = a1, s2 = a2, f0 = 0, f1 = 0, f2 = 0;
        for (int j0 = 0; j0 < nev; j0 += kNevReg) {
            if (nev > kNevReg) {   // multi-chunk: stream the coefficients again (L2-resident)
#pragma unroll
                for (int j = 0; j < kNevReg; j++) {
                    ev[j][0] = ev[j][1] = ev[j][2] = 0.0;
                    if (on && j0 + j < nev) {
                        const double* c = EV + (size_t)(j0 + j) * 3 * S;
                        ev[j][0] = c[tid]; ev[j][1] = c[S + tid]; ev[j][2] = c[2 * S + tid];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kNevReg; j++) {
                if (j0 + j < nev) {
                    const double p = sProj[j0 + j], w = sPw[j0 + j];
                    s0 = fma(ev[j][0], p, s0); s1 = fma(ev[j][1], p, s1); s2 = fma(ev[j][2], p, s2);
                    f0 = fma(-ev[j][0], w, f0); f1 = fma(-ev[j][1], w, f1); f2 = fma(-ev[j][2], w, f2);
                }
            }
        }
        // back to the laboratory frame (edmd.cpp:130-150): x = R^T (s - shift), F = R^T f
        if (on) {
            s0 -= shift0; s1 -= shift1; s2 -= shift2;
            X[tid]         = R[0] * s0 + R[3] * s1 + R[6] * s2;
            X[S + tid]     = R[1] * s0 + R[4] * s1 + R[7] * s2;
            X[2 * S + tid] = R[2] * s0 + R[5] * s1 + R[8] * s2;
            double* F = A.fed + (size_t)t * 3 * S;
            F[tid]         = R[0] * f0 + R[3] * f1 + R[6] * f2;
            F[S + tid]     = R[1] * f0 + R[4] * f1 + R[7] * f2;
            F[2 * S + tid] = R[2] * f0 + R[5] * f1 + R[8] * f2;
        }
        if (tid == 0) {   // edmd.cpp:153-157
            double en = 0.0;
            for (int j = 0; j < nev; j++) en += sProj[j] * sProj[j] / EL[j];
            A.e_ed[t] = en * (0.5 * A.scaled);
        }
    }
}

// ---- warp-per-tetrad projection (parameter sets with <= kNevReg eigenvectors) -----------------------
// The CTA stages the parameter set's eigenvectors [12][3][S] and average structure [3][S] in shared
// memory once per run of tetrads that share the set (`order` is sorted by set); after that every
// warp projects its own tetrads with no CTA barrier: lane l owns atoms l, l+32, ... (8 per lane),
// the twelve partial projections are summed with the halving butterfly and handed back to all lanes
// with shuffles.  ~2,000 warp instructions per tetrad against ~6,800 for the thread-per-atom kernel
// (whose eight warps each run the butterfly and then meet at two barriers per tetrad).
// Bound: shared-memory bandwidth, 2 passes x 36 coefficients x 256 atoms x 8 B = 147 KB per tetrad.
__device__ __forceinline__ constexpr int reduce12_owner(int j) {   // inverse of reduce12_index (even lane)
    return 16 * (j / 6) + 8 * ((j % 6) / 3) + ((j % 3) == 2 ? 4 : 0) + ((j % 3) == 1 ? 2 : 0);
}

template <int SC>   // SC: the atom stride when it is the usual 256 (all shared-memory offsets fold into immediates), else 0
__global__ void __launch_bounds__(kThreads, 2) ed_project_warp_kernel(EdArgs A) {
    extern __shared__ double wsm[];
    __shared__ int sNext;
    const int S = SC ? SC : A.ps.S;
    double* sEV = wsm;               // [kNevReg][3][S]
    double* sAV = wsm + kNevReg * 3 * S;   // [3][S]
    const int tid = threadIdx.x, lane = tid & 31;
    // warp-uniform values pass through redux.sync (result in a uniform register): the compiler then sees the
    // loops over tetrads as warp-uniform and drops the reconvergence code around every shuffle in them
    const int warp = __reduce_max_sync(0xffffffffu, tid >> 5);
    const int e0 = blockIdx.x * A.chunk;
    const int e1 = (e0 + A.chunk < A.count) ? e0 + A.chunk : A.count;
    const int nk = S >> 5;
    const int my_idx = reduce12_index(lane);

    int base = e0;
    while (base < e1) {
        const int ps = __reduce_max_sync(0xffffffffu, A.param_id[A.order[base]]);
        const int na = __reduce_max_sync(0xffffffffu, A.ps.na[ps]), nev = __reduce_max_sync(0xffffffffu, A.ps.nev[ps]);
        if (tid == 0) sNext = e1;
        {   // stage the set's tables (rows j >= NEV are zero; the arrays are zero-padded beyond na / nev)
            const double* EV = A.ps.evec + (size_t)ps * A.ps.NEV * 3 * S;
            const double* AV = A.ps.avg + (size_t)ps * 3 * S;
            const int have = A.ps.NEV * 3 * S;
            for (int i = tid; i < kNevReg * 3 * S; i += kThreads) sEV[i] = i < have ? EV[i] : 0.0;
            for (int i = tid; i < 3 * S; i += kThreads) sAV[i] = AV[i];
        }
        __syncthreads();
        double el = 1.0;
        if (my_idx >= 0 && my_idx < nev) el = A.ps.eval[(size_t)ps * A.ps.NEV + my_idx];

        int e = base + warp;
        int t = __reduce_max_sync(0xffffffffu, e < e1 ? A.order[e] : -1);
        while (e < e1) {
            if (__reduce_max_sync(0xffffffffu, A.param_id[t]) != ps) break;
            const int e_next = e + kWarps;
            const int t_next = __reduce_max_sync(0xffffffffu, e_next < e1 ? A.order[e_next] : -1);
            const double* XIN = A.crd_in + (size_t)t * 3 * S;
            const double* SUP = A.sup + (size_t)t * kSupWords;
            double x[8][3];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int a = lane + 32 * k;
                x[k][0] = x[k][1] = x[k][2] = 0.0;
                if (k < nk && a < na) { x[k][0] = XIN[a]; x[k][1] = XIN[S + a]; x[k][2] = XIN[2 * S + a]; }
            }
            double v[kNevReg];
#pragma unroll
            for (int j = 0; j < kNevReg; j++) v[j] = 0.0;
            {
                double R[9];
#pragma unroll
                for (int q = 0; q < 9; q++) R[q] = SUP[q];
                const double ca0 = SUP[12], ca1 = SUP[13], ca2 = SUP[14];
                const double cx0 = SUP[15], cx1 = SUP[16], cx2 = SUP[17];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int a = lane + 32 * k;
                    if (k < nk) {
                        double d0 = 0, d1 = 0, d2 = 0;
                        if (a < na) {   // edmd.cpp:89-94 (centred copies as qcprot.c:382-386)
                            const double ac0 = sAV[a] - ca0, ac1 = sAV[S + a] - ca1, ac2 = sAV[2 * S + a] - ca2;
                            const double xc0 = x[k][0] - cx0, xc1 = x[k][1] - cx1, xc2 = x[k][2] - cx2;
                            d0 = R[0] * xc0 + R[1] * xc1 + R[2] * xc2 - ac0;
                            d1 = R[3] * xc0 + R[4] * xc1 + R[5] * xc2 - ac1;
                            d2 = R[6] * xc0 + R[7] * xc1 + R[8] * xc2 - ac2;
                        }
#pragma unroll
                        for (int j = 0; j < kNevReg; j++) {   // edmd.cpp:106-110
                            const double* c = sEV + (size_t)j * 3 * S + a;
                            v[j] = fma(c[0], d0, fma(c[S], d1, fma(c[2 * S], d2, v[j])));
                        }
                    }
                }
            }
            // warp totals; the lane that owns projection j derives the force weight and energy term
            const double mine = warp_reduce12(v, lane);
            const double pw_mine = mine * A.scaled / el;
            const double en_mine = mine * mine / el;
            double p[kNevReg], w[kNevReg];
            double en = 0.0;
#pragma unroll
            for (int j = 0; j < kNevReg; j++) {
                p[j] = __shfl_sync(0xffffffffu, mine, reduce12_owner(j));
                w[j] = __shfl_sync(0xffffffffu, pw_mine, reduce12_owner(j));
                const double q = __shfl_sync(0xffffffffu, en_mine, reduce12_owner(j));
                if (j < nev) en += q;                          // edmd.cpp:153-157
            }
            if (lane == 0) A.e_ed[t] = en * (0.5 * A.scaled);

            // re-embedding, force (edmd.cpp:113-127) and back to the laboratory frame (:130-150)
            double R[9];
#pragma unroll
            for (int q = 0; q < 9; q++) R[q] = SUP[q];
            const double shift0 = SUP[9], shift1 = SUP[10], shift2 = SUP[11];
            double* X = A.crd + (size_t)t * 3 * S;
            double* F = A.fed + (size_t)t * 3 * S;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int a = lane + 32 * k;
                if (k < nk && a < na) {
                    double s0 = sAV[a], s1 = sAV[S + a], s2 = sAV[2 * S + a], f0 = 0, f1 = 0, f2 = 0;
#pragma unroll
                    for (int j = 0; j < kNevReg; j++) {
                        const double* c = sEV + (size_t)j * 3 * S + a;
                        const double c0 = c[0], c1 = c[S], c2 = c[2 * S];
                        s0 = fma(c0, p[j], s0); s1 = fma(c1, p[j], s1); s2 = fma(c2, p[j], s2);
                        f0 = fma(-c0, w[j], f0); f1 = fma(-c1, w[j], f1); f2 = fma(-c2, w[j], f2);
                    }
                    s0 -= shift0; s1 -= shift1; s2 -= shift2;
                    X[a]         = R[0] * s0 + R[3] * s1 + R[6] * s2;
                    X[S + a]     = R[1] * s0 + R[4] * s1 + R[7] * s2;
                    X[2 * S + a] = R[2] * s0 + R[5] * s1 + R[8] * s2;
                    F[a]         = R[0] * f0 + R[3] * f1 + R[6] * f2;
                    F[S + a]     = R[1] * f0 + R[4] * f1 + R[7] * f2;
                    F[2 * S + a] = R[2] * f0 + R[5] * f1 + R[8] * f2;
                }
            }
            // (Writing the distance pass's bounding spheres here, while the coordinates are in L1/L2, was tried:
            // the epilogue cost 0.18 ms at 1e5 tetrads — what the separate group_bounds_kernel costs — because this
            // kernel runs 16 warps per SM and the sphere code is a chain of dependent loads.  Dropped.)
            e = e_next; t = t_next;
        }
        if (lane == 0 && e < e1) atomicMin(&sNext, e);   // first tetrad of the next set seen by this warp
        __syncthreads();
        base = __reduce_max_sync(0xffffffffu, sNext);
        __syncthreads();
    }
}

cudaError_t launch_superpose(const ParamSets& ps, const int* param_id, const double* crd_in, double* sup, int t0, int count,
                             int variant, int sms, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    SupArgs B{ps, param_id, crd_in, sup, t0, count};
    // a few thousand tetrads do not fill the machine either way: then the shorter per-warp critical path of the
    // one-tetrad kernel wins (1,000 tetrads: 17 vs 30 us); beyond that the two-tetrad kernel's throughput does
    if (variant == 0) variant = count < 4096 ? 1 : 2;
    if (variant == 1) superpose_kernel<<<(count + kSupWarps - 1) / kSupWarps, 32 * kSupWarps, 0, st>>>(B);
    else {
        // (fewer CTAs per SM — tried by padding the CTA with unused shared memory — is monotonically slower:
        // 1.15 / 1.16 / 1.20 / 1.21 / 1.31 ms of ED at 8 / 7 / 6 / 5 / 4 CTAs per SM: the kernel lives on occupancy)
        // One tetrad pair per warp.  (A "balanced" grid — every warp the same number of pairs, no thin last wave — was
        // measured SLOWER at 1e5 and at 12,500 tetrads: short CTAs handed out by the hardware scheduler beat long ones.)
        const int grid = ((count + 1) / 2 + kS2Warps - 1) / kS2Warps;
        (void)sms;
        if (ps.S == kMaxStride) superpose2_kernel<kMaxStride><<<grid, 32 * kS2Warps, 0, st>>>(B);
        else                    superpose2_kernel<0><<<grid, 32 * kS2Warps, 0, st>>>(B);
    }
    return cudaGetLastError();
}

cudaError_t launch_ed_project(const ParamSets& ps, const int* param_id, const double* crd_in, double* crd, double* fed,
                              double* e_ed, double* sup, const int* order, int sms, int t0, int count, double scaled,
                              bool use_warp_kernel, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    cudaError_t err;
    if (use_warp_kernel && ps.NEV <= kNevReg) {
        const size_t smem = sizeof(double) * (kNevReg * 3 + 3) * ps.S;
        // the opt-in to > 48 KB of dynamic shared memory is per device: remember which devices have it
        static bool done[64] = {};
        int dev = 0;
        err = cudaGetDevice(&dev);
        if (err != cudaSuccess) return err;
        bool& configured = done[dev & 63];
        if (!configured) {
            const int most = (int)(sizeof(double) * (kNevReg * 3 + 3) * kMaxStride);
            err = cudaFuncSetAttribute(ed_project_warp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
            if (err != cudaSuccess) return err;
            err = cudaFuncSetAttribute(ed_project_warp_kernel<kMaxStride>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
            if (err != cudaSuccess) return err;
            configured = true;
        }
        int chunk = (count + 2 * sms - 1) / (2 * sms);   // two resident CTAs per SM, one wave
        if (chunk < kWarps) chunk = kWarps;              // at least one tetrad per warp
        EdArgs A{ps, param_id, crd_in, crd, fed, e_ed, sup, t0, scaled, order, count, chunk};
        if (ps.S == kMaxStride) ed_project_warp_kernel<kMaxStride><<<(count + chunk - 1) / chunk, kThreads, smem, st>>>(A);
        else                    ed_project_warp_kernel<0><<<(count + chunk - 1) / chunk, kThreads, smem, st>>>(A);
        return cudaGetLastError();
    }
    int chunk = count / (sms * 8);          // keep >= 8 CTAs per SM before making runs longer
    if (chunk < 1) chunk = 1;
    if (chunk > 8) chunk = 8;
    EdArgs A{ps, param_id, crd_in, crd, fed, e_ed, sup, t0, scaled, order, count, chunk};
    ed_project_kernel<<<(count + chunk - 1) / chunk, kThreads, 0, st>>>(A);
    return cudaGetLastError();
}

}  // namespace edmd
