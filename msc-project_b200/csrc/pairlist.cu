// pairlist.cu — tetrad centroids and the centroid-cutoff pair list.
//
// Replaces Master::cal_Centre_of_Mass (/root/reference/src/master.cpp:157-176) and
// Master::generate_Pair_Lists (:180-210): same inclusion predicate
//     |c_i - c_j|^2 < mole_Cutoff^2  &&  |i-j| > mole_Least  &&  |i-j| < n - mole_Least,
// same orientation rule (odd separation -> (j,i), even -> (i,j)) and the same order
// (i ascending, then j ascending), so per-tetrad force sums run in the reference's order.
// The pair count is 64-bit (the reference's int overflows at n >= 46,341, master.cpp:77).
//
// The inclusion test is discontinuous, so the centroid is summed sequentially in atom order
// with the reference's expression tree (no FMA): borderline pairs fall on the same side as on
// the CPU.  Runs every ntsync steps only.
#include "common.cuh"
#include "kernels.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace edmd {

__global__ void centroid_kernel(const double* crd, const int* natoms, double* cen, int n, int S) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double* X = crd + (size_t)t * 3 * S;
    const int na = natoms[t];
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int a = 0; a < na; a++) {
        cx = __dadd_rn(cx, X[a]); cy = __dadd_rn(cy, X[S + a]); cz = __dadd_rn(cz, X[2 * S + a]);
    }
    cen[4 * (size_t)t] = cx / na; cen[4 * (size_t)t + 1] = cy / na; cen[4 * (size_t)t + 2] = cz / na;
    cen[4 * (size_t)t + 3] = 0.0;
}

__device__ __forceinline__ bool pair_included(const double* cen, int i, int j, int n, double cut2,
                                              double least) {
    const double dx = __dsub_rn(cen[4 * (size_t)i], cen[4 * (size_t)j]);
    const double dy = __dsub_rn(cen[4 * (size_t)i + 1], cen[4 * (size_t)j + 1]);
    const double dz = __dsub_rn(cen[4 * (size_t)i + 2], cen[4 * (size_t)j + 2]);
    const double r = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const int sep = j - i;
    return (r < cut2) && ((double)sep > least) && ((double)sep < ((double)n - least));
}

// One warp per row i; lanes sweep j = i+1 .. n-1 in ascending chunks of 32.
// pass 0: count; pass 1: ordered fill at row_off[i].
__global__ void pair_rows_kernel(const double* cen, int n, double cut2, double least,
                                 long long* row_count, const long long* row_off, int2* pairs, int fill) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_block) {
        long long cnt = 0;
        const long long off = fill ? row_off[i] : 0;
        for (int j0 = i + 1; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool in = (j < n) && pair_included(cen, i, j, n, cut2, least);
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (fill && in) {
                const int rank = __popc(bal & ((1u << lane) - 1u));
                const int sep = j - i;
                pairs[off + cnt + rank] = (sep & 1) ? make_int2(j, i) : make_int2(i, j);   // master.cpp:197-201
            }
            cnt += __popc(bal);
        }
        if (!fill && lane == 0) row_count[i] = cnt;
    }
}

// ---- cell list ---------------------------------------------------------------------------------
// The O(n^2) sweep above is the reference's algorithm (master.cpp:188-189); beyond ~1e5 tetrads it is
// replaced by a uniform grid of cells of edge mole_Cutoff: tetrads are radix-sorted by cell id
// (stable, so ascending inside a cell), a warp per row visits the 27 neighbouring cells through
// binary searches in the sorted cell-id array (no dense cell table, any extent), and each row's
// partners are sorted ascending afterwards — the SAME list in the SAME order as the sweep.
struct CellGrid {
    double x0, y0, z0, inv_h;
    long long nx, ny, nz;
};

__device__ __forceinline__ long long cell_of(const CellGrid& G, const double* c, long long& ix, long long& iy, long long& iz) {
    ix = (long long)floor((c[0] - G.x0) * G.inv_h);
    iy = (long long)floor((c[1] - G.y0) * G.inv_h);
    iz = (long long)floor((c[2] - G.z0) * G.inv_h);
    ix = ix < 0 ? 0 : (ix >= G.nx ? G.nx - 1 : ix);
    iy = iy < 0 ? 0 : (iy >= G.ny ? G.ny - 1 : iy);
    iz = iz < 0 ? 0 : (iz >= G.nz ? G.nz - 1 : iz);
    return ix + G.nx * (iy + G.ny * iz);
}

__global__ void cell_id_kernel(const double* cen, int n, CellGrid G, unsigned long long* cell, int* ident) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    long long ix, iy, iz;
    cell[t] = (unsigned long long)cell_of(G, cen + 4 * (size_t)t, ix, iy, iz);
    ident[t] = t;
}

__device__ __forceinline__ int lower_bound_u64(const unsigned long long* a, int n, unsigned long long key) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// fill == 0: row_count[i] = number of partners j > i.  fill == 1: partners written (unordered) at
// row_off[i].  One warp per row.
__global__ void cell_rows_kernel(const double* cen, int n, double cut2, double least, CellGrid G,
                                 const unsigned long long* sorted_cell, const int* sorted_tet,
                                 long long* row_count, const long long* row_off, int* partners, int fill) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_block) {
        long long ix, iy, iz;
        cell_of(G, cen + 4 * (size_t)i, ix, iy, iz);
        int lo = 0, hi = 0;
        if (lane < 27) {
            const long long cx = ix + (lane % 3) - 1, cy = iy + ((lane / 3) % 3) - 1, cz = iz + (lane / 9) - 1;
            if (cx >= 0 && cx < G.nx && cy >= 0 && cy < G.ny && cz >= 0 && cz < G.nz) {
                const unsigned long long c = (unsigned long long)(cx + G.nx * (cy + G.ny * cz));
                lo = lower_bound_u64(sorted_cell, n, c);
                hi = lower_bound_u64(sorted_cell, n, c + 1);
            }
        }
        long long cnt = 0;
        const long long off = fill ? row_off[i] : 0;
        for (int k = 0; k < 27; k++) {
            const int a = __shfl_sync(0xffffffffu, lo, k), b = __shfl_sync(0xffffffffu, hi, k);
            for (int m0 = a; m0 < b; m0 += 32) {
                const int m = m0 + lane;
                int j = -1;
                bool in = false;
                if (m < b) { j = sorted_tet[m]; in = (j > i) && pair_included(cen, i, j, n, cut2, least); }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (fill && in) partners[off + cnt + __popc(bal & ((1u << lane) - 1u))] = j;
                cnt += __popc(bal);
            }
        }
        if (!fill && lane == 0) row_count[i] = cnt;
    }
}

// rows of sorted partners -> (t1,t2) with the orientation rule of master.cpp:197-201
__global__ void rows_to_pairs_kernel(const int* partners, const long long* row_off, int n, int2* pairs) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_block) {
        for (long long p = row_off[i] + lane; p < row_off[i + 1]; p += 32) {
            const int j = partners[p];
            pairs[p] = ((j - i) & 1) ? make_int2(j, i) : make_int2(i, j);
        }
    }
}

cudaError_t cell_sort_tetrads(const double* cen, int n, const double lo[3], const double hi[3], double h,
                              CellGridHost* grid_out, unsigned long long* cell, unsigned long long* cell_sorted,
                              int* tet, int* tet_sorted, void** tmp, size_t* tmp_bytes, cudaStream_t st) {
    CellGrid G;
    G.x0 = lo[0]; G.y0 = lo[1]; G.z0 = lo[2]; G.inv_h = 1.0 / h;
    G.nx = (long long)floor((hi[0] - lo[0]) / h) + 1; G.ny = (long long)floor((hi[1] - lo[1]) / h) + 1;
    G.nz = (long long)floor((hi[2] - lo[2]) / h) + 1;
    grid_out->x0 = G.x0; grid_out->y0 = G.y0; grid_out->z0 = G.z0; grid_out->inv_h = G.inv_h;
    grid_out->nx = G.nx; grid_out->ny = G.ny; grid_out->nz = G.nz;
    cell_id_kernel<<<(n + 255) / 256, 256, 0, st>>>(cen, n, G, cell, tet);
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, cell, cell_sorted, tet, tet_sorted, n, 0, 64, st);
    if (need > *tmp_bytes) {
        cudaFree(*tmp);
        cudaError_t err = cudaMalloc(tmp, need);
        if (err != cudaSuccess) { *tmp = nullptr; *tmp_bytes = 0; return err; }
        *tmp_bytes = need;
    }
    return cub::DeviceRadixSort::SortPairs(*tmp, need, cell, cell_sorted, tet, tet_sorted, n, 0, 64, st);
}

cudaError_t launch_cell_rows(const double* cen, int n, double cut2, double least, const CellGridHost& g,
                             const unsigned long long* cell_sorted, const int* tet_sorted, long long* row_count,
                             const long long* row_off, int* partners, int fill, cudaStream_t st) {
    CellGrid G{g.x0, g.y0, g.z0, g.inv_h, g.nx, g.ny, g.nz};
    const int wpb = 8;
    int grid = (n + wpb - 1) / wpb;
    if (grid > 148 * 16) grid = 148 * 16;
    cell_rows_kernel<<<grid, wpb * 32, 0, st>>>(cen, n, cut2, least, G, cell_sorted, tet_sorted, row_count, row_off,
                                                partners, fill);
    return cudaGetLastError();
}

cudaError_t sort_rows_and_emit(const int* partners, int* partners_sorted, const long long* row_off, int n,
                               long long np, int2* pairs, void** tmp, size_t* tmp_bytes, cudaStream_t st) {
    if (np > 0) {
        size_t need = 0;
        cub::DeviceSegmentedSort::SortKeys(nullptr, need, partners, partners_sorted, np, n, row_off, row_off + 1, st);
        if (need > *tmp_bytes) {
            cudaFree(*tmp);
            cudaError_t err = cudaMalloc(tmp, need);
            if (err != cudaSuccess) { *tmp = nullptr; *tmp_bytes = 0; return err; }
            *tmp_bytes = need;
        }
        cudaError_t err = cub::DeviceSegmentedSort::SortKeys(*tmp, need, partners, partners_sorted, np, n, row_off,
                                                             row_off + 1, st);
        if (err != cudaSuccess) return err;
    }
    const int wpb = 8;
    int grid = (n + wpb - 1) / wpb;
    if (grid > 148 * 16) grid = 148 * 16;
    rows_to_pairs_kernel<<<grid, wpb * 32, 0, st>>>(partners_sorted, row_off, n, pairs);
    return cudaGetLastError();
}

// ---- device-side bookkeeping of a rebuild (nothing but three scalars crosses PCIe) -------------------
static cudaError_t grow_tmp(void** tmp, size_t* tmp_bytes, size_t need) {
    if (need <= *tmp_bytes) return cudaSuccess;
    cudaFree(*tmp);
    cudaError_t err = cudaMalloc(tmp, need);
    if (err != cudaSuccess) { *tmp = nullptr; *tmp_bytes = 0; return err; }
    *tmp_bytes = need;
    return cudaSuccess;
}

// Bounding box of the centroids: out = {lo x,y,z, hi x,y,z, all finite ? 1 : 0}.  One CTA; n <= a few 1e6.
__global__ void __launch_bounds__(1024) centroid_bbox_kernel(const double* cen, int n, double* out) {
    __shared__ double slo[3][32], shi[3][32];
    __shared__ int sbad[32];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    int bad = 0;
    for (int t = threadIdx.x; t < n; t += blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double v = cen[4 * (size_t)t + c];
            if (!(v > -1e15 && v < 1e15)) bad = 1;
            lo[c] = fmin(lo[c], v); hi[c] = fmax(hi[c], v);
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if (lane == 0) { for (int c = 0; c < 3; c++) { slo[c][warp] = lo[c]; shi[c][warp] = hi[c]; } sbad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; w++) {
            for (int c = 0; c < 3; c++) { lo[c] = fmin(lo[c], slo[c][w]); hi[c] = fmax(hi[c], shi[c][w]); }
            bad |= sbad[w];
        }
        for (int c = 0; c < 3; c++) { out[c] = lo[c]; out[3 + c] = hi[c]; }
        out[6] = bad ? 0.0 : 1.0;
    }
}
cudaError_t launch_centroid_bbox(const double* cen, int n, double* out7, cudaStream_t st) {
    centroid_bbox_kernel<<<1, 1024, 0, st>>>(cen, n, out7);
    return cudaGetLastError();
}

// row_off[0] = 0, row_off[i+1] = sum of row_count[0..i]
cudaError_t scan_row_counts(const long long* row_count, long long* row_off, int n, void** tmp, size_t* tmp_bytes,
                            cudaStream_t st) {
    cudaError_t err = cudaMemsetAsync(row_off, 0, sizeof(long long), st);
    if (err != cudaSuccess) return err;
    size_t need = 0;
    cub::DeviceScan::InclusiveSum(nullptr, need, row_count, row_off + 1, n, st);
    if ((err = grow_tmp(tmp, tmp_bytes, need)) != cudaSuccess) return err;
    return cub::DeviceScan::InclusiveSum(*tmp, need, row_count, row_off + 1, n, st);
}

// CSR over tetrads of the pair list, entries of a tetrad in pair-list order (worker.cpp:166-179 accumulates
// in that order): every pair contributes two (tetrad, 2*pair + side) records, a STABLE radix sort by tetrad
// keeps each tetrad's records in ascending pair order.
__global__ void adj_records_kernel(const int2* pairs, long long np, int* key, unsigned long long* val) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
        const int2 pr = pairs[p];
        key[2 * p] = pr.x;     val[2 * p] = (unsigned long long)p << 1;
        key[2 * p + 1] = pr.y; val[2 * p + 1] = ((unsigned long long)p << 1) | 1ull;
    }
}
__global__ void adj_emit_kernel(const unsigned long long* val_sorted, const int2* pairs, long long total,
                                int* adj_partner, int* adj_pair) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const unsigned long long v = val_sorted[e];
        const long long p = (long long)(v >> 1);
        const int2 pr = pairs[p];
        adj_pair[e] = (int)p;
        adj_partner[e] = (v & 1ull) ? (pr.x | (int)0x80000000) : pr.y;   // bit 31: this tetrad is t2 of the pair
    }
}
__global__ void adj_offsets_kernel(const int* key_sorted, long long total, int n, long long* adj_off) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n) return;
    long long lo = 0, hi = total;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (key_sorted[mid] < t) lo = mid + 1; else hi = mid; }
    adj_off[t] = lo;
}
cudaError_t build_adjacency_device(const int2* pairs, long long np, int n, int* key, int* key_sorted,
                                   unsigned long long* val, unsigned long long* val_sorted, long long* adj_off,
                                   int* adj_partner, int* adj_pair, void** tmp, size_t* tmp_bytes, cudaStream_t st) {
    const long long total = 2 * np;
    if (np > 0) {
        long long blocks = (np + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        adj_records_kernel<<<(int)blocks, 256, 0, st>>>(pairs, np, key, val);
        int bits = 1;
        while ((1LL << bits) < (long long)n) bits++;
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, key, key_sorted, val, val_sorted, total, 0, bits, st);
        cudaError_t err = grow_tmp(tmp, tmp_bytes, need);
        if (err != cudaSuccess) return err;
        err = cub::DeviceRadixSort::SortPairs(*tmp, need, key, key_sorted, val, val_sorted, total, 0, bits, st);
        if (err != cudaSuccess) return err;
        long long eb = (total + 255) / 256;
        if (eb > 148 * 16) eb = 148 * 16;
        adj_emit_kernel<<<(int)eb, 256, 0, st>>>(val_sorted, pairs, total, adj_partner, adj_pair);
    }
    adj_offsets_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(key_sorted, total, n, adj_off);
    return cudaGetLastError();
}

cudaError_t launch_centroids(const double* crd, const int* natoms, double* cen, int n, int S, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    centroid_kernel<<<(n + 127) / 128, 128, 0, st>>>(crd, natoms, cen, n, S);
    return cudaGetLastError();
}

cudaError_t launch_pair_rows(const double* cen, int n, double cut2, double least, long long* row_count,
                             const long long* row_off, int2* pairs, int fill, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int wpb = 8;
    int grid = (n + wpb - 1) / wpb;
    if (grid > 148 * 16) grid = 148 * 16;
    pair_rows_kernel<<<grid, wpb * 32, 0, st>>>(cen, n, cut2, least, row_count, row_off, pairs, fill);
    return cudaGetLastError();
}

}  // namespace edmd
