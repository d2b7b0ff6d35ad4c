// step_math.cuh — per-tetrad velocity/coordinate update shared by integrate_kernel and the fused
// post-scan kernel (nb.cu).  Replaces EDMD::update_Velocities (/root/reference/src/edmd.cpp:289-317)
// and EDMD::update_Coordinates (:322-327).
#pragma once
#include "common.cuh"

namespace edmd {

struct IntArgs {
    ParamSets ps;
    const int* param_id;
    double* vel;          // [n][3][S] in/out
    const double* crd;    // [n][3][S] coordinates in
    double* crd_out;      // [n][3][S] coordinates out (may alias crd when no other CTA reads crd)
    const double* fed;    // [n][3][S]
    const double* rnd;    // [n][3][S] or nullptr (= zero random terms)
    const double* fnb;    // [n][3][S] (ignored by the fused kernel, which has F_nb in registers)
    double* temp;         // [n] out
    int t0;
    int do_vel, do_crd;
    double dt, gamma, tautp, scaled;
};

// This thread's operands of the update, loaded up front so that their HBM latency overlaps whatever
// the caller does next (the fused kernel walks the pair adjacency in between).
struct IntLoad {
    double v[3], fe[3], r[3], m[3], x[3];
    int na;
    bool on;
};

__device__ __forceinline__ void integrate_load(const IntArgs& A, int t, IntLoad& L) {
    const int tid = threadIdx.x;
    const int ps = A.param_id[t];
    const int S = A.ps.S;
    L.na = A.ps.na[ps];
    L.on = tid < L.na;
    const size_t base = (size_t)t * 3 * S;
#pragma unroll
    for (int c = 0; c < 3; c++) { L.v[c] = L.fe[c] = L.r[c] = L.x[c] = 0.0; L.m[c] = 1.0; }
    if (L.on) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const size_t k = base + c * S + tid;
            L.v[c] = A.vel[k];
            if (A.do_vel) {
                L.fe[c] = A.fed[k];
                L.r[c] = A.rnd ? A.rnd[k] : 0.0;
                L.m[c] = A.ps.mass[(size_t)ps * 3 * S + c * S + tid];
            }
            if (A.do_crd) L.x[c] = A.crd[k];
        }
    }
}

// One CTA = one tetrad, thread tid <-> atom tid.  fnb_reg: this atom's clipped NB force.
__device__ __forceinline__ void integrate_apply(const IntArgs& A, int t, IntLoad& L, const double fnb_reg[3],
                                                double (*red)[kWarps]) {
    const int tid = threadIdx.x;
    const int S = A.ps.S;
    const int na = L.na;
    const size_t base = (size_t)t * 3 * S;
    if (A.do_vel) {
        const double gamfac = 1.0 / (1.0 + A.gamma * A.dt);
        double ke[1] = {0.0};
        if (L.on) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                // edmd.cpp:299 — note the ED force is not divided by the mass
                L.v[c] = (L.v[c] + L.fe[c] * A.dt + (L.r[c] + fnb_reg[c]) * A.dt / L.m[c]) * gamfac;
                ke[0] += 0.5 * L.m[c] * L.v[c] * L.v[c];
            }
        }
        block_sum<1>(ke, red);
        const double target = 0.5 * A.scaled * 3 * na;                              // :305
        const double tscal = sqrt(1.0 + (A.dt / A.tautp) * ((target / ke[0]) - 1.0)); // :306
        if (tid == 0) A.temp[t] = ke[0] * 2 / (0.002 * 3 * na) * (tscal * tscal);      // :309-310
        if (L.on) {
#pragma unroll
            for (int c = 0; c < 3; c++) { L.v[c] *= tscal; A.vel[base + c * S + tid] = L.v[c]; }
        }
    }
    if (A.do_crd && L.on) {
#pragma unroll
        for (int c = 0; c < 3; c++) A.crd_out[base + c * S + tid] = L.x[c] + L.v[c] * A.dt;   // :325
    }
}

__device__ __forceinline__ void integrate_tetrad(const IntArgs& A, int t, const double fnb_reg[3],
                                                 double (*red)[kWarps]) {
    IntLoad L;
    integrate_load(A, t, L);
    integrate_apply(A, t, L, fnb_reg, red);
}

}  // namespace edmd
