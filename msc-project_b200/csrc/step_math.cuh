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

// One CTA = one tetrad, thread tid <-> atom tid.  fnb_reg: this atom's clipped NB force.
__device__ __forceinline__ void integrate_tetrad(const IntArgs& A, int t, const double fnb_reg[3],
                                                 double (*red)[kWarps]) {
    const int tid = threadIdx.x;
    const int ps = A.param_id[t];
    const int na = A.ps.na[ps];
    const int S = A.ps.S;
    const bool on = tid < na;
    const size_t base = (size_t)t * 3 * S;
    double v[3] = {0, 0, 0};
    if (on) {
#pragma unroll
        for (int c = 0; c < 3; c++) v[c] = A.vel[base + c * S + tid];
    }
    if (A.do_vel) {
        const double gamfac = 1.0 / (1.0 + A.gamma * A.dt);
        double ke[1] = {0.0};
        if (on) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const size_t k = base + c * S + tid;
                const double m = A.ps.mass[(size_t)ps * 3 * S + c * S + tid];
                const double r = A.rnd ? A.rnd[k] : 0.0;
                // edmd.cpp:299 — note the ED force is not divided by the mass
                v[c] = (v[c] + A.fed[k] * A.dt + (r + fnb_reg[c]) * A.dt / m) * gamfac;
                ke[0] += 0.5 * m * v[c] * v[c];
            }
        }
        block_sum<1>(ke, red);
        const double target = 0.5 * A.scaled * 3 * na;                              // :305
        const double tscal = sqrt(1.0 + (A.dt / A.tautp) * ((target / ke[0]) - 1.0)); // :306
        if (tid == 0) A.temp[t] = ke[0] * 2 / (0.002 * 3 * na) * (tscal * tscal);      // :309-310
        if (on) {
#pragma unroll
            for (int c = 0; c < 3; c++) { v[c] *= tscal; A.vel[base + c * S + tid] = v[c]; }
        }
    }
    if (A.do_crd && on) {
#pragma unroll
        for (int c = 0; c < 3; c++) A.crd_out[base + c * S + tid] = A.crd[base + c * S + tid] + v[c] * A.dt;  // :325
    }
}

}  // namespace edmd
