// kernels.h — host-callable launchers of the EDMD device kernels (one per .cu file).
#pragma once
#include <cuda_runtime.h>
#include <string.h>
#include "common.cuh"

namespace edmd {

// Atom pairs with r^2 >= this cannot change any force or energy (see nb.cu header).
inline double nb_effective_cut2(const SimParams& sp) {
    const double cut2 = sp.atom_cutoff * sp.atom_cutoff;
    if (sp.qfac == 0.0) return cut2 < 2.0 ? cut2 : 2.0;
    return cut2;
}

constexpr int kSupWordsHost = 24;   // doubles per tetrad of superposition scratch (ed.cu)
// variant: 0 = by size, 1 = one tetrad per warp with the coordinates staged in shared memory, 2 = two tetrads per warp
cudaError_t launch_superpose(const ParamSets& ps, const int* param_id, const double* crd_in, double* sup, int t0, int count,
                             int variant, int sms, cudaStream_t st);
cudaError_t launch_ed_project(const ParamSets& ps, const int* param_id, const double* crd_in, double* crd, double* fed,
                              double* e_ed, double* sup, const int* order, int sms, int t0, int count, double scaled,
                              bool use_warp_kernel, cudaStream_t st);
// the projection as DMMA contractions over groups of eight tetrads that share a parameter set (ed_mma.cu)
bool ed_project_mma_supported(const ParamSets& ps);
cudaError_t launch_ed_project_mma(const ParamSets& ps, const int* param_id, const double* crd_in, double* crd, double* fed,
                                  double* e_ed, const double* sup, const int* order, int sms, int count, double scaled,
                                  cudaStream_t st);

cudaError_t launch_nb_scan(const double* crd, const int* natoms, const int2* pairs, unsigned long long* hit,
                           long long p0, long long p1, int S, double cut2_eff, int mode, int sms,
                           unsigned long long* const* peers, int world, cudaStream_t st);

cudaError_t launch_nb_scan_cull(const double* crd, const int* natoms, const int2* pairs, unsigned long long* hit,
                                long long p0, long long p1, int n, int S, double cut2_eff, int sms, double* gb,
                                unsigned long long* items, long long items_cap, unsigned* item_counts /* [2] */,
                                int dense_min, const int* param_id, const int* perm, const int* need_list, int need_count,
                                unsigned long long* const* peers, int world, cudaStream_t st);
cudaError_t launch_mark_needed(const int2* pairs, long long p0, long long p1, unsigned char* need, int* list, int* count,
                               int n, cudaStream_t st);

// hit: per-pair 8 x 8 block masks written by the distance pass (0 = pair not flagged); perm / grp: the
// per-parameter-set slot -> atom and atom -> slot-group tables the masks refer to (engine.cu); epart: [n][8][2]
// per-slot-group pair energies; ticket: work-item counter
cudaError_t launch_nb_accumulate(const double* crd, const int* natoms, const int* param_id, const double* chg,
                                 const unsigned long long* hit, const long long* adj_off, const int* adj_partner,
                                 const int* adj_pair, const int* perm, const unsigned long long* grp, double* fnb,
                                 double* e_nb, double* epart, unsigned* ticket, int t0, int count, int S, int sms,
                                 const SimParams& sp, int clip, cudaStream_t st);

// mark: int [2n + 4] scratch (flags, list counter, work-item counter, distance-pass item counter, pad, compact list
// of marked tetrads); bump: the random-stream call counter to advance by bump_by (nullptr: none)
cudaError_t launch_nb_finish_split(const double* crd_ro, const int* natoms, const int* param_id, const double* chg,
                                   const unsigned long long* hit, const int2* pairs, long long np, const long long* adj_off,
                                   const int* adj_partner, const int* adj_pair, const int* perm,
                                   const unsigned long long* grp, double* fnb, double* e_nb, double* epart,
                                   const ParamSets& ps, double* vel, double* crd_out, const double* fed,
                                   const double* rnd, double* temp, int* mark, unsigned char* dirty, int n,
                                   int t0, int count, int S, int sms, const SimParams& sp, long* bump, long bump_by,
                                   cudaStream_t st);

cudaError_t launch_integrate(const ParamSets& ps, const int* param_id, double* vel, double* crd,
                             const double* fed, const double* rnd, const double* fnb, double* temp,
                             int t0, int count, int do_vel, int do_crd, const SimParams& sp,
                             cudaStream_t st);

cudaError_t launch_merge(double* vel, double* crd, const int* bp_off, int n, int S, cudaStream_t st);

cudaError_t launch_centroids(const double* crd, const int* natoms, double* cen, int n, int S, cudaStream_t st);
cudaError_t launch_pair_rows(const double* cen, int n, double cut2, double least, long long* row_count,
                             const long long* row_off, int2* pairs, int fill, cudaStream_t st);

// cell-list pair builder (pairlist.cu)
struct CellGridHost { double x0, y0, z0, inv_h; long long nx, ny, nz; };
cudaError_t cell_sort_tetrads(const double* cen, int n, const double lo[3], const double hi[3], double h,
                              CellGridHost* grid_out, unsigned long long* cell, unsigned long long* cell_sorted,
                              int* tet, int* tet_sorted, void** tmp, size_t* tmp_bytes, cudaStream_t st);
cudaError_t launch_cell_rows(const double* cen, int n, double cut2, double least, const CellGridHost& g,
                             const unsigned long long* cell_sorted, const int* tet_sorted, long long* row_count,
                             const long long* row_off, int* partners, int fill, cudaStream_t st);
cudaError_t sort_rows_and_emit(const int* partners, int* partners_sorted, const long long* row_off, int n,
                               long long np, int2* pairs, void** tmp, size_t* tmp_bytes, cudaStream_t st);

// device-side bookkeeping of a pair-list rebuild (pairlist.cu)
cudaError_t launch_centroid_bbox(const double* cen, int n, double* out7, cudaStream_t st);
cudaError_t scan_row_counts(const long long* row_count, long long* row_off, int n, void** tmp, size_t* tmp_bytes,
                            cudaStream_t st);
cudaError_t build_adjacency_device(const int2* pairs, long long np, int n, int* key, int* key_sorted,
                                   unsigned long long* val, unsigned long long* val_sorted, long long* adj_off,
                                   int* adj_partner, int* adj_pair, void** tmp, size_t* tmp_bytes, cudaStream_t st);

// multi-GPU exchange over peer memory (peer.cu)
cudaError_t launch_peer_barrier(unsigned long long* const* flags, unsigned long long* epoch, int* err, int world,
                                int rank, unsigned long long timeout_ns, cudaStream_t st);
cudaError_t launch_halo_pull(double* local, double* const* peer, const int* list, int count, int chunk, int S, int sms,
                             cudaStream_t st);
cudaError_t build_halo_list(const int2* pairs, long long np, long long p0, long long p1, int t0, int t1, int n,
                            unsigned char* flag, int* list, int* count, cudaStream_t st);

// glibc-rand()/Kinderman-Monahan random terms, one warp per tetrad (rng.cu)
cudaError_t launch_random_terms(const ParamSets& ps, const double* noise, const int* param_id, double* rnd, int t0, int count,
                                const SimParams& sp, long time0, int rank, long calls_before,
                                const long* calls_dev, int max_ctas /* 0: one tetrad per warp; > 0: at most that many CTAs */, cudaStream_t st);
cudaError_t launch_ps_precompute(const ParamSets& ps, int P, double* pre, cudaStream_t st);
cudaError_t launch_noise_table(const ParamSets& ps, int P, double* noise, const SimParams& sp, cudaStream_t st);
cudaError_t launch_bump_counter(long* counter, long by, cudaStream_t st);

// interleaved xyz (reference layout, tetrads concatenated at off3[t]) <-> SoA planes
cudaError_t launch_pack_planes(const double* flat, const long long* off3, const int* natoms, double* planes,
                               int n, int S, cudaStream_t st);
cudaError_t launch_unpack_planes(const double* planes, const long long* off3, const int* natoms, double* flat,
                                 int n, int S, cudaStream_t st);

// DFMA throughput microbenchmark (peak.cu)
cudaError_t measure_fp64_peak(double* tflops, double* sm_mhz_est);

}  // namespace edmd
