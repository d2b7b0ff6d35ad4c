// bounds.cuh — bounding spheres of one tetrad's slot groups and of the whole tetrad (culled distance pass, nb.cu).
// Shared by group_bounds_kernel and the epilogue of the ED projection (ed.cu), which has just produced the
// coordinates.  One warp per tetrad.  Conservative by construction: radii are reduced in FP32 with round-up
// conversions and inflated, so a sphere always contains its atoms whatever the rounding of the distances.
#pragma once
#include "common.cuh"

namespace edmd {

// X: the tetrad's planes [3][S]; PM: its parameter set's slot -> atom table [kMaxStride] (-1 = empty slot, slot
// 32g = group g's 1-centre); gb: [n][8][4] group spheres followed by [n][4] tetrad spheres (x, y, z, radius; -1 = empty)
__device__ __forceinline__ void group_bounds_tetrad(const double* X, const int* PM, double* gb, int n, int t, int S,
                                                    int lane) {
    const unsigned full = 0xffffffffu;
    int a[8];
#pragma unroll
    for (int g = 0; g < 8; g++) a[g] = PM[32 * g + lane];     // slot -> atom, -1 = empty slot
    double gx = 0.0, gy = 0.0, gz = 0.0, gr = -1.0;            // lane g keeps group g's sphere
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const bool on = a[g] >= 0;
        const double x = on ? X[a[g]] : 0.0, y = on ? X[S + a[g]] : 0.0, z = on ? X[2 * S + a[g]] : 0.0;
        const double cx = __shfl_sync(full, x, 0), cy = __shfl_sync(full, y, 0), cz = __shfl_sync(full, z, 0);
        const bool any = __shfl_sync(full, (int)on, 0) != 0;   // slots fill from 0: empty slot 0 = empty group
        const double d2 = on ? (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz) : 0.0;
        float f = __double2float_ru(d2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) f = fmaxf(f, __shfl_xor_sync(full, f, o));
        const double rad = any ? (double)__fsqrt_ru(f) * (1.0 + 1e-12) + 1e-12 : -1.0;
        if (lane == 0) *reinterpret_cast<double4*>(gb + ((size_t)t * 8 + g) * 4) = make_double4(cx, cy, cz, rad);
        if (lane == g) { gx = cx; gy = cy; gz = cz; gr = rad; }
    }
    // tetrad sphere: centre = mean of the non-empty group centres, radius = max over groups of
    // (distance to the centre + group radius); lanes 0..7 hold the groups
    const bool have = lane < 8 && gr >= 0.0;
    double tx = have ? gx : 0.0, ty = have ? gy : 0.0, tz = have ? gz : 0.0; int tn = have ? 1 : 0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        tx += __shfl_xor_sync(full, tx, o); ty += __shfl_xor_sync(full, ty, o);
        tz += __shfl_xor_sync(full, tz, o); tn += __shfl_xor_sync(full, tn, o);
    }
    // (every lane runs the shuffles below: tn is non-zero in lanes 0..7 only, and a shuffle inside a
    // branch that part of the named lanes skip is undefined)
    const double inv = tn > 0 ? (double)__frcp_rn((float)tn) : 0.0;   // any centre will do: the radius is measured from it
    tx *= inv; ty *= inv; tz *= inv;
    const double dd = (gx - tx) * (gx - tx) + (gy - ty) * (gy - ty) + (gz - tz) * (gz - tz);
    float reach = have ? __fadd_ru(__fsqrt_ru(__double2float_ru(dd)), __double2float_ru(gr)) : 0.0f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) reach = fmaxf(reach, __shfl_xor_sync(full, reach, o));
    const double trad = tn > 0 ? (double)reach * (1.0 + 1e-12) + 1e-12 : -1.0;
    if (lane == 0) *reinterpret_cast<double4*>(gb + (size_t)n * 32 + (size_t)t * 4) = make_double4(tx, ty, tz, trad);
}

}  // namespace edmd
