// ed.cu — ED restoring force: QCP superposition + projection on the PCA eigenvectors.
//
// Replaces EDMD::calculate_ED_Forces (/root/reference/src/edmd.cpp:64-166) and the QCP
// routines it calls (src/qcprot/qcprot.c:90-407), batched over tetrads.
//
// One CTA of 256 threads per tetrad, thread a <-> atom a.  Each thread keeps its atom's
// 3 x nev eigenvector coefficients in registers between the projection pass and the
// re-embedding pass, so the 72.6 KB eigenvector block crosses the memory system once
// (the reference walks it twice, the second time with stride 3N, edmd.cpp:106-127).
// Reductions (centroids, 3x3 cross-covariance, projections) are fixed-order CTA sums.
// Roofline: HBM — 96,872 algorithmic bytes per tetrad (SURVEY.md §8a a-ED); with
// de-duplicated parameter sets resident in L2 the DRAM traffic drops to the 24 KB of state.
#include "common.cuh"
#include "kernels.h"

namespace edmd {

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

struct EdArgs {
    ParamSets ps;
    const int* param_id;
    double* crd;        // [n][3][S] in/out (overwritten with the re-embedded coordinates)
    double* fed;        // [n][3][S] out
    double* e_ed;       // [n] out
    int t0;             // first tetrad of this launch
    double scaled;
};

__global__ void __launch_bounds__(kThreads, 2) ed_forces_kernel(EdArgs A) {
    __shared__ double red[16][kWarps];
    __shared__ double sR[9];
    __shared__ double sProj[64];
    __shared__ double sPw[64];

    const int t = A.t0 + blockIdx.x;
    const int tid = threadIdx.x;
    const int ps = A.param_id[t];
    const int na = A.ps.na[ps], nev = A.ps.nev[ps];
    const int S = A.ps.S;
    const bool on = tid < na;
    double* X = A.crd + (size_t)t * 3 * S;
    const double* AV = A.ps.avg + (size_t)ps * 3 * S;
    const double* EV = A.ps.evec + (size_t)ps * A.ps.NEV * 3 * S;
    const double* EL = A.ps.eval + (size_t)ps * A.ps.NEV;

    double x0 = 0, x1 = 0, x2 = 0, a0 = 0, a1 = 0, a2 = 0;
    if (on) {
        x0 = X[tid]; x1 = X[S + tid]; x2 = X[2 * S + tid];
        a0 = AV[tid]; a1 = AV[S + tid]; a2 = AV[2 * S + tid];
    }
    // first eigenvector chunk: issue the loads early, they are independent of the QCP
    double ev[kNevReg][3];
    const int nchunk0 = nev < kNevReg ? nev : kNevReg;
#pragma unroll
    for (int j = 0; j < kNevReg; j++) {
        ev[j][0] = ev[j][1] = ev[j][2] = 0.0;
        if (on && j < nchunk0) {
            const double* e = EV + (size_t)j * 3 * S;
            ev[j][0] = e[tid]; ev[j][1] = e[S + tid]; ev[j][2] = e[2 * S + tid];
        }
    }

    // centroids — CenterCoords, qcprot.c:370-387
    double c[6] = {a0, a1, a2, x0, x1, x2};
    block_sum<6>(c, red);
    double ac0 = 0, ac1 = 0, ac2 = 0, xc0 = 0, xc1 = 0, xc2 = 0;
    if (on) {
        ac0 = a0 - c[0] / na; ac1 = a1 - c[1] / na; ac2 = a2 - c[2] / na;
        xc0 = x0 - c[3] / na; xc1 = x1 - c[4] / na; xc2 = x2 - c[5] / na;
    }
    // cross-covariance + inner products — InnerProduct, qcprot.c:132-160
    double m[11] = {ac0 * xc0, ac0 * xc1, ac0 * xc2, ac1 * xc0, ac1 * xc1, ac1 * xc2,
                    ac2 * xc0, ac2 * xc1, ac2 * xc2,
                    ac0 * ac0 + ac1 * ac1 + ac2 * ac2, xc0 * xc0 + xc1 * xc1 + xc2 * xc2};
    block_sum<11>(m, red);
    if (tid < 32) {   // one warp solves the quartic; the other seven wait at the barrier
        double R[9];
        qcp_rotation(R, m, (m[9] + m[10]) * 0.5);
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 9; k++) sR[k] = R[k];
        }
    }
    __syncthreads();
    double R[9];
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = sR[k];

    // displacement in the reference frame (edmd.cpp:89-94) and offset vector (:97-102)
    double d0 = 0, d1 = 0, d2 = 0, h0 = 0, h1 = 0, h2 = 0;
    if (on) {
        d0 = R[0] * xc0 + R[1] * xc1 + R[2] * xc2 - ac0;
        d1 = R[3] * xc0 + R[4] * xc1 + R[5] * xc2 - ac1;
        d2 = R[6] * xc0 + R[7] * xc1 + R[8] * xc2 - ac2;
        h0 = a0 - (R[0] * x0 + R[1] * x1 + R[2] * x2);
        h1 = a1 - (R[3] * x0 + R[4] * x1 + R[5] * x2);
        h2 = a2 - (R[6] * x0 + R[7] * x1 + R[8] * x2);
    }

    // projections (edmd.cpp:106-110), kNevReg eigenvectors per pass
    double shift0 = 0, shift1 = 0, shift2 = 0;
    for (int j0 = 0; j0 < nev || j0 == 0; j0 += kNevReg) {
        if (j0 > 0) {
#pragma unroll
            for (int j = 0; j < kNevReg; j++) {
                ev[j][0] = ev[j][1] = ev[j][2] = 0.0;
                if (on && j0 + j < nev) {
                    const double* e = EV + (size_t)(j0 + j) * 3 * S;
                    ev[j][0] = e[tid]; ev[j][1] = e[S + tid]; ev[j][2] = e[2 * S + tid];
                }
            }
        }
        double v[kNevReg + 3];
#pragma unroll
        for (int j = 0; j < kNevReg; j++) v[j] = ev[j][0] * d0 + ev[j][1] * d1 + ev[j][2] * d2;
        v[kNevReg] = h0; v[kNevReg + 1] = h1; v[kNevReg + 2] = h2;
        block_sum<kNevReg + 3>(v, red);
        if (j0 == 0) { shift0 = v[kNevReg] / na; shift1 = v[kNevReg + 1] / na; shift2 = v[kNevReg + 2] / na; }
        if (tid < kNevReg && j0 + tid < nev) {
            sProj[j0 + tid] = v[tid];
            sPw[j0 + tid] = v[tid] * A.scaled / EL[j0 + tid];
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
                    const double* e = EV + (size_t)(j0 + j) * 3 * S;
                    ev[j][0] = e[tid]; ev[j][1] = e[S + tid]; ev[j][2] = e[2 * S + tid];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kNevReg; j++) {
            if (j0 + j < nev) {
                const double p = sProj[j0 + j], w = sPw[j0 + j];
                s0 += ev[j][0] * p; s1 += ev[j][1] * p; s2 += ev[j][2] * p;
                f0 -= ev[j][0] * w; f1 -= ev[j][1] * w; f2 -= ev[j][2] * w;
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
        double e = 0.0;
        for (int j = 0; j < nev; j++) e += sProj[j] * sProj[j] / EL[j];
        A.e_ed[t] = e * (0.5 * A.scaled);
    }
}

cudaError_t launch_ed_forces(const ParamSets& ps, const int* param_id, double* crd, double* fed,
                             double* e_ed, int t0, int count, double scaled, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    EdArgs A{ps, param_id, crd, fed, e_ed, t0, scaled};
    ed_forces_kernel<<<count, kThreads, 0, st>>>(A);
    return cudaGetLastError();
}

}  // namespace edmd
