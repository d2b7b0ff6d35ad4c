// engine.cu — the C ABI (include/edmd_b200.h) over the sm_100a kernels.
//
// Host-side replacement for the reference's orchestration of the hot path: Master/Worker
// message protocol (/root/reference/src/master.cpp:268-343, src/worker.cpp:147-199) becomes
// in-order kernel launches on one CUDA stream; MPI derived datatypes (src/mpilib.cpp) become
// device-resident SoA planes.  No CPU fallback exists: every compute entry point launches
// CUDA kernels or fails.
#include "../../include/edmd_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace edmd;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess)                                                            \
            return fail(EDMD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                              \
    } while (0)

enum TimerId { T_ED = 0, T_SCAN, T_ACC, T_RNG, T_MERGE, T_PAIRS, T_INT, T_LAYOUT, T_SUP, T_COUNT };

struct Timer {
    std::vector<cudaEvent_t> ev;   // start/stop pairs not yet folded
    double total_ms = 0.0;
    long long launches = 0;
};

template <class T>
cudaError_t dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    return cudaMalloc((void**)p, count * sizeof(T));
}

}  // namespace

struct edmd_engine {
    int device = 0;
    int sms = 148;
    cudaStream_t st = nullptr;
    SimParams sp{};

    int n = 0, S = 0, P = 0, NEV = 0;
    long long total3 = 0;
    std::vector<int> natoms_h, nev_h, param_id_h;
    std::vector<long long> off3_h;

    // device: metadata + parameter sets
    int *d_natoms = nullptr, *d_param_id = nullptr, *d_ps_na = nullptr, *d_ps_nev = nullptr;
    int* d_order = nullptr;   // owned tetrads [t0,t1) sorted by parameter set (ED project kernel)
    long long* d_off3 = nullptr;
    double *d_avg = nullptr, *d_mass = nullptr, *d_chg = nullptr, *d_eval = nullptr, *d_evec = nullptr;
    ParamSets ps{};

    // device: state planes
    double *crd = nullptr, *vel = nullptr, *fed = nullptr, *rnd = nullptr, *fnb = nullptr;
    double* crd2 = nullptr;        // post-integrate coordinates written by the fused finish kernel
    bool latest_in_crd2 = false;   // owned tetrads' newest coordinates live in crd2 (see normalize())
    double* obs = nullptr;         // one block [4n]: e_ed [n] | e_nb [n][2] | temp [n] (one peer mapping serves all three)
    double *e_ed = nullptr, *e_nb = nullptr, *temp = nullptr, *cen = nullptr, *sup = nullptr;
    int* d_mark = nullptr;            // [2n+1] per-tetrad "in a flagged pair" flags, list counter, list of marked tetrads
    unsigned char* d_dirty = nullptr; // [n] fnb of this tetrad is non-zero from an earlier step
    double* d_epart = nullptr;        // [n][8][2] pair energies per slot group (force pass -> finish kernel)
    double* d_gb = nullptr;                 // [n][8][4] group + [n][4] tetrad bounding spheres (scan mode 4)
    int* d_perm = nullptr;                  // [P][kMaxStride] slot -> atom, compact groups of 32 (culling scan, force pass)
    unsigned long long* d_grp = nullptr;    // [P][32] atom -> slot group, byte c of entry l = group of atom 32c+l
    unsigned long long* d_pmask = nullptr;  // counter + [pairs_cap] {pair index, 8x8 group mask} work items (scan mode 4)
    long long pmask_cap = 0;
    bool rnd_zero = true;   // random terms currently all zero

    // staging (interleaved flat), 3 slots
    double* d_stage = nullptr;
    double* h_stage = nullptr;   // pinned

    // pair list + adjacency
    int2* d_pairs = nullptr;
    long long np = 0, pairs_cap = 0;
    unsigned long long* d_hit = nullptr;   // [pairs_cap] per-pair block masks of the distance pass (0 = not flagged)
    long long *d_row_count = nullptr, *d_row_off = nullptr, *d_adj_off = nullptr;
    int *d_adj_partner = nullptr, *d_adj_pair = nullptr;
    long long adj_cap = 0;
    bool have_pairs = false;
    bool hits_valid = false;

    // cell-list builder scratch (pairlist.cu), grown on demand
    unsigned long long *d_cell = nullptr, *d_cell_sorted = nullptr;
    int *d_tet = nullptr, *d_tet_sorted = nullptr, *d_partners = nullptr, *d_partners_sorted = nullptr;
    void* d_cub_tmp = nullptr; size_t cub_tmp_bytes = 0; long long partners_cap = 0;
    int pair_builder = 0;   // 0 auto, 1 sweep (reference O(n^2)), 2 cell list

    // device-side CSR builder scratch (pairlist.cu: build_adjacency_device), grown on demand
    int *d_adj_key = nullptr, *d_adj_key_sorted = nullptr;
    unsigned long long *d_adj_val = nullptr, *d_adj_val_sorted = nullptr;
    long long adj_rec_cap = 0;
    double* d_bbox = nullptr;   // [7] centroid bounding box + finite flag

    // ---- multi-GPU over peer memory (peer.cu) --------------------------------------------------------
    // Every rank maps four buffers of every other rank: coordinates, velocities, the per-pair block masks
    // the distance kernels publish into, the barrier flags, and the per-tetrad observables.  Between
    // processes the mappings are CUDA IPC handles (edmd_peer_export / edmd_peer_attach); inside one process
    // (edmd_group_*) they are the raw pointers with peer access enabled.
    unsigned long long* d_hitx = nullptr;   // peer-visible mask buffer (replaces d_hit while peers are attached)
    long long hitx_cap = 0;
    unsigned long long* d_flags = nullptr;  // [32] arrival counters written by the peers' barrier kernels
    unsigned long long* d_epoch = nullptr;  // barrier number
    int* d_peer_err = nullptr;              // set by a barrier that timed out
    double** d_crd_tbl = nullptr;           // device arrays [world]: every rank's buffer as mapped here
    unsigned long long** d_hit_tbl = nullptr;
    unsigned long long** d_flag_tbl = nullptr;
    std::vector<double*> peer_crd, peer_vel, peer_obs;   // host copies of the mapped pointers (own rank: local)
    std::vector<void*> peer_opened;         // IPC mappings to close
    int world = 1, rank_id = 0;
    bool peer_mode = false;
    bool peer_replicated = false;           // every rank runs the per-tetrad kernels for all tetrads (small systems)
    int* d_halo = nullptr; int* d_halo_count = nullptr; unsigned char* d_halo_flag = nullptr;
    int halo_count = 0;                     // non-owned tetrads read by this rank's step
    unsigned long long peer_timeout_ns = 60ull * 1000000000ull;

    // merge
    int* d_bp_off = nullptr;
    bool have_bp = false;

    // shard
    int t0 = 0, t1 = 0;
    long long p0 = 0, p1 = 0;
    bool shard_pairs_custom = false;

    // rng stream position
    long time0 = 1000000000L;
    int rank = 1;
    long calls = 0;

    int ed_kernel = 0;   // projection kernel: 0 automatic, 1 thread-per-atom, 2 warp-per-tetrad, 3 tensor-core (edmd_set_ed_kernel)
    bool ed_auto_mma = true;    // "automatic" picks the tensor-core projection for large systems with shared sets (EDMD_ED_MMA=0: never)
    int cull_dense_min = 9;   // sphere masks with this many blocks or more go to the tiled distance kernel (EDMD_CULL_DENSE)
    int sup_variant = 0; // superposition kernel: 0 by size, 1 one tetrad per warp, 2 two tetrads per warp (EDMD_SUPERPOSE)
    int scan_mode = 4;   // 0 direct, 1 gram, 2 fp32 prefilter, 4 gram + culling (default; see nb.cu)
    // CUDA graph of one step (edmd_step replays it; rebuilt whenever anything it bakes in changes)
    bool use_graph = true, graph_valid = false;
    int graph_random_mode = -1;
    cudaGraph_t graph = nullptr; cudaGraphExec_t graph_exec = nullptr;
    long* d_calls = nullptr;   // device copy of the random-stream call counter
    int nb_clip = 1;     // clip summed NB force components to [-1,1] (master.cpp:304-305)
    cudaEvent_t span0 = nullptr, span1 = nullptr;   // edmd_timer_start/stop
    std::vector<cudaEvent_t> marks;                 // edmd_mark / edmd_mark_elapsed
    bool timing = false;
    Timer timers[T_COUNT];
    long long launches = 0;
    long long graph_launches = 0;   // kernel launches inside one replay of the step graph
    unsigned char* d_need = nullptr; bool need_valid = false;   // [n] tetrad is an end of a pair in [p0,p1) (scan mode 4)
    int* d_need_list = nullptr; int need_count = 0;             // the same tetrads as a compact list (+ its device counter)
    double* d_pspre = nullptr;              // [P][4] per-set superposition constants (ed.cu)
    double* d_noise = nullptr; bool noise_valid = false;   // [P][3][S] Langevin noise amplitudes
    cudaStream_t st_copy = nullptr;                 // edmd_step_host: velocity upload beside the ED / scan kernels
    cudaEvent_t ev_crd_in = nullptr, ev_vel_in = nullptr;
    cudaStream_t st_rng = nullptr;                  // edmd_step: the random-term generator beside the ED / distance kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool rng_branch = false;   // measured slower than back-to-back kernels: off unless asked for
    int rng_ctas_per_sm = 16;
    // asynchronous output (edmd_snapshot_begin / _wait): device -> caller buffers on a side stream
    cudaStream_t st_out = nullptr;
    cudaEvent_t ev_staged = nullptr, ev_snap[2] = {nullptr, nullptr};
    std::atomic<bool> snap_pending[2] = {{false}, {false}};   // cleared by edmd_snapshot_wait, possibly on a writer thread
    double* d_obs_snap = nullptr;                   // [4n] copy of the observables taken in stream order
};

namespace {

void timer_begin(edmd_engine* e, int id) {
    if (!e->timing) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, e->st);
    e->timers[id].ev.push_back(a);
    e->timers[id].ev.push_back(b);
}
void timer_end(edmd_engine* e, int id, int launches) {
    e->launches += launches;
    if (!e->timing) return;
    Timer& t = e->timers[id];
    cudaEventRecord(t.ev.back(), e->st);
    t.launches += launches;
}
void timer_fold(edmd_engine* e, int id) {
    Timer& t = e->timers[id];
    if (t.ev.empty()) return;
    cudaStreamSynchronize(e->st);
    for (size_t i = 0; i + 1 < t.ev.size(); i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.ev[i], t.ev[i + 1]) == cudaSuccess) t.total_ms += ms;
        cudaEventDestroy(t.ev[i]); cudaEventDestroy(t.ev[i + 1]);
    }
    t.ev.clear();
}

void drop_graph(edmd_engine* e);
void close_peers(edmd_engine* e);
void free_system(edmd_engine* e) {
    drop_graph(e);
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
    close_peers(e);
    cudaFree(e->d_hitx); e->d_hitx = nullptr; e->hitx_cap = 0;
    cudaFree(e->d_flags); cudaFree(e->d_epoch); cudaFree(e->d_peer_err);
    e->d_flags = e->d_epoch = nullptr; e->d_peer_err = nullptr;
    cudaFree(e->d_halo); cudaFree(e->d_halo_count); cudaFree(e->d_halo_flag);
    e->d_halo = e->d_halo_count = nullptr; e->d_halo_flag = nullptr; e->halo_count = 0;
    cudaFree(e->d_adj_key); cudaFree(e->d_adj_key_sorted); cudaFree(e->d_adj_val); cudaFree(e->d_adj_val_sorted);
    e->d_adj_key = e->d_adj_key_sorted = nullptr; e->d_adj_val = e->d_adj_val_sorted = nullptr; e->adj_rec_cap = 0;
    cudaFree(e->d_bbox); e->d_bbox = nullptr;
    cudaFree(e->d_obs_snap); e->d_obs_snap = nullptr;
    e->snap_pending[0] = e->snap_pending[1] = false;
    cudaFree(e->d_calls); e->d_calls = nullptr;
    cudaFree(e->d_mark); cudaFree(e->d_dirty); e->d_mark = nullptr; e->d_dirty = nullptr;
    cudaFree(e->d_epart); e->d_epart = nullptr;
    cudaFree(e->d_perm); e->d_perm = nullptr;
    cudaFree(e->d_grp); e->d_grp = nullptr;
    cudaFree(e->d_gb); cudaFree(e->d_pmask); e->d_gb = nullptr; e->d_pmask = nullptr; e->pmask_cap = 0;
    cudaFree(e->d_noise); e->d_noise = nullptr; e->noise_valid = false;
    cudaFree(e->d_pspre); e->d_pspre = nullptr;
    cudaFree(e->d_need); e->d_need = nullptr; e->need_valid = false;
    cudaFree(e->d_need_list); e->d_need_list = nullptr; e->need_count = 0;
    cudaFree(e->d_order); e->d_order = nullptr;
    cudaFree(e->d_cell); cudaFree(e->d_cell_sorted); cudaFree(e->d_tet); cudaFree(e->d_tet_sorted);
    cudaFree(e->d_partners); cudaFree(e->d_partners_sorted); cudaFree(e->d_cub_tmp);
    e->d_cell = e->d_cell_sorted = nullptr; e->d_tet = e->d_tet_sorted = e->d_partners = e->d_partners_sorted = nullptr;
    e->d_cub_tmp = nullptr; e->cub_tmp_bytes = 0; e->partners_cap = 0;
    cudaFree(e->d_natoms); cudaFree(e->d_param_id); cudaFree(e->d_ps_na); cudaFree(e->d_ps_nev);
    cudaFree(e->d_off3); cudaFree(e->d_avg); cudaFree(e->d_mass); cudaFree(e->d_chg);
    cudaFree(e->d_eval); cudaFree(e->d_evec);
    cudaFree(e->crd); cudaFree(e->crd2); cudaFree(e->vel); cudaFree(e->fed); cudaFree(e->rnd); cudaFree(e->fnb);
    cudaFree(e->obs); e->obs = nullptr; cudaFree(e->cen); cudaFree(e->sup);
    cudaFree(e->d_stage); if (e->h_stage) cudaFreeHost(e->h_stage);
    cudaFree(e->d_pairs); cudaFree(e->d_hit); cudaFree(e->d_row_count); cudaFree(e->d_row_off);
    cudaFree(e->d_adj_off); cudaFree(e->d_adj_partner); cudaFree(e->d_adj_pair); cudaFree(e->d_bp_off);
    e->d_natoms = e->d_param_id = e->d_ps_na = e->d_ps_nev = nullptr; e->d_off3 = nullptr;
    e->d_avg = e->d_mass = e->d_chg = e->d_eval = e->d_evec = nullptr;
    e->crd2 = nullptr; e->latest_in_crd2 = false;
    e->crd = e->vel = e->fed = e->rnd = e->fnb = e->e_ed = e->e_nb = e->temp = e->cen = e->sup = nullptr;
    e->d_stage = nullptr; e->h_stage = nullptr;
    e->d_pairs = nullptr; e->d_hit = nullptr; e->d_row_count = e->d_row_off = e->d_adj_off = nullptr;
    e->d_adj_partner = e->d_adj_pair = nullptr; e->d_bp_off = nullptr;
    e->np = e->pairs_cap = e->adj_cap = 0; e->have_pairs = e->hits_valid = e->have_bp = false;
    e->n = 0;
}

uint64_t hash_bytes(uint64_t h, const void* p, size_t bytes) {
    const uint64_t* w = (const uint64_t*)p;
    for (size_t i = 0; i < bytes / 8; i++) { h ^= w[i]; h *= 0x100000001b3ULL; h ^= h >> 29; }
    return h;
}

struct PtrKey {
    const void *a, *m, *q, *l, *v; int na, nev;
    bool operator==(const PtrKey& o) const {
        return a == o.a && m == o.m && q == o.q && l == o.l && v == o.v && na == o.na && nev == o.nev;
    }
};
struct PtrKeyHash {
    size_t operator()(const PtrKey& k) const {
        return (size_t)((uintptr_t)k.a * 31 + (uintptr_t)k.v * 17 + (uintptr_t)k.m * 7 + (uintptr_t)k.l + k.na);
    }
};

int ensure_pair_capacity(edmd_engine* e, long long np) {
    if (e->peer_mode && np > e->hitx_cap)
        return fail(EDMD_ERR_STATE, "pair list of %lld pairs outgrew the peer-visible mask buffer (%lld): "
                    "edmd_peer_detach on every rank, then export and attach again", np, e->hitx_cap);
    if (np <= e->pairs_cap) return 0;
    long long cap = std::max(np, e->pairs_cap * 3 / 2 + 1024);
    cudaFree(e->d_pairs); cudaFree(e->d_hit); cudaFree(e->d_pmask);
    e->d_pairs = nullptr; e->d_hit = nullptr; e->d_pmask = nullptr; e->pairs_cap = 0; e->pmask_cap = 0;
    CK(dev_alloc(&e->d_pairs, (size_t)cap));
    CK(dev_alloc(&e->d_hit, (size_t)cap));
    CK(dev_alloc(&e->d_pmask, 2 * (size_t)cap + 2));   // one {pair, mask} work item per pair (nb.cu)
    e->pmask_cap = cap;
    e->pairs_cap = cap;
    return 0;
}

// CSR over tetrads of the pair list, entries in pair-list order (worker.cpp:166-179 accumulates in that
// order), built on the device from e->d_pairs (pairlist.cu).
int build_adjacency(edmd_engine* e) {
    const long long np = e->np;
    if (np >= (1LL << 30)) return fail(EDMD_ERR_LIMIT, "%lld pairs exceeds the 2^30 limit of the adjacency builder", np);
    if (2 * np > e->adj_cap) {
        cudaFree(e->d_adj_partner); cudaFree(e->d_adj_pair);
        e->d_adj_partner = e->d_adj_pair = nullptr; e->adj_cap = 0;
        const long long cap = 2 * np + 1024;
        CK(dev_alloc(&e->d_adj_partner, (size_t)cap));
        CK(dev_alloc(&e->d_adj_pair, (size_t)cap));
        e->adj_cap = cap;
    }
    if (2 * np > e->adj_rec_cap) {
        cudaFree(e->d_adj_key); cudaFree(e->d_adj_key_sorted); cudaFree(e->d_adj_val); cudaFree(e->d_adj_val_sorted);
        e->d_adj_key = e->d_adj_key_sorted = nullptr; e->d_adj_val = e->d_adj_val_sorted = nullptr; e->adj_rec_cap = 0;
        const long long cap = 2 * np + np / 2 + 1024;
        CK(dev_alloc(&e->d_adj_key, (size_t)cap)); CK(dev_alloc(&e->d_adj_key_sorted, (size_t)cap));
        CK(dev_alloc(&e->d_adj_val, (size_t)cap)); CK(dev_alloc(&e->d_adj_val_sorted, (size_t)cap));
        e->adj_rec_cap = cap;
    }
    CK(build_adjacency_device(e->d_pairs, np, e->n, e->d_adj_key, e->d_adj_key_sorted, e->d_adj_val, e->d_adj_val_sorted,
                              e->d_adj_off, e->d_adj_partner, e->d_adj_pair, &e->d_cub_tmp, &e->cub_tmp_bytes, e->st));
    e->launches += np > 0 ? 4 : 1;
    return 0;
}

int check_ready(edmd_engine* e, bool need_pairs) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    if (e->n == 0) return fail(EDMD_ERR_STATE, "no tetrads uploaded (call edmd_upload_tetrads first)");
    if (need_pairs && !e->have_pairs)
        return fail(EDMD_ERR_STATE, "no pair list (call edmd_build_pairlist or edmd_set_pairlist first)");
    return 0;
}

void close_peers(edmd_engine* e) {
    for (void* p : e->peer_opened) cudaIpcCloseMemHandle(p);
    e->peer_opened.clear();
    cudaFree(e->d_crd_tbl); cudaFree(e->d_hit_tbl); cudaFree(e->d_flag_tbl);
    e->d_crd_tbl = nullptr; e->d_hit_tbl = nullptr; e->d_flag_tbl = nullptr;
    e->peer_crd.clear(); e->peer_vel.clear(); e->peer_obs.clear();
    e->peer_mode = false; e->world = 1; e->rank_id = 0; e->halo_count = 0;
}

unsigned long long* cur_hits(edmd_engine* e) { return e->peer_mode ? e->d_hitx : e->d_hit; }

// flag barrier across the attached ranks, in stream order (peer.cu)
int peer_barrier(edmd_engine* e) {
    CK(launch_peer_barrier(e->d_flag_tbl, e->d_epoch, e->d_peer_err, e->world, e->rank_id, e->peer_timeout_ns, e->st));
    e->launches++;
    return 0;
}
int peer_check(edmd_engine* e) {   // after a synchronisation: did a barrier give up?
    if (!e->peer_mode || !e->d_peer_err) return 0;
    int bad = 0;
    CK(cudaMemcpy(&bad, e->d_peer_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) return fail(EDMD_ERR_STATE, "a peer barrier timed out (rank %d of %d): a peer stopped stepping", e->rank_id, e->world);
    return 0;
}
// the tetrads [a,b) that rank r owns: contiguous blocks of ceil(n/world), or everything when replicated
void owned_range(const edmd_engine* e, int r, int* a, int* b) {
    if (e->peer_replicated) { *a = 0; *b = e->n; return; }
    const int chunk = (e->n + e->world - 1) / e->world;
    *a = std::min(e->n, r * chunk); *b = std::min(e->n, *a + chunk);
}
// which non-owned tetrads a step of this rank reads (peer.cu); called whenever the shard or the list changes
int rebuild_halo(edmd_engine* e) {
    e->halo_count = 0;
    if (!e->peer_mode || e->peer_replicated || !e->have_pairs) return 0;
    if (!e->d_halo) {
        CK(dev_alloc(&e->d_halo, (size_t)e->n)); CK(dev_alloc(&e->d_halo_count, 1)); CK(dev_alloc(&e->d_halo_flag, (size_t)e->n));
    }
    CK(build_halo_list(e->d_pairs, e->np, e->p0, e->p1, e->t0, e->t1, e->n, e->d_halo_flag, e->d_halo, e->d_halo_count, e->st));
    CK(cudaMemcpyAsync(&e->halo_count, e->d_halo_count, sizeof(int), cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    e->launches += 2;
    return 0;
}

void drop_graph(edmd_engine* e) {
    e->need_valid = false;   // every caller changes the pair list, the shard or the scan mode
    // the executable graph is kept: the next capture tries to update it in place (edmd_step)
    if (e->graph) { cudaGraphDestroy(e->graph); e->graph = nullptr; }
    e->graph_valid = false;
}

// Owned tetrads sorted (stably) by parameter set: consecutive tetrads of a CTA's run share the
// eigenvector registers in ed_project_kernel.
int upload_order(edmd_engine* e) {
    std::vector<int> ord;
    ord.reserve(e->t1 - e->t0);
    for (int t = e->t0; t < e->t1; t++) ord.push_back(t);
    std::stable_sort(ord.begin(), ord.end(), [e](int a, int b) { return e->param_id_h[a] < e->param_id_h[b]; });
    if (!e->d_order) CK(dev_alloc(&e->d_order, (size_t)e->n));
    if (!ord.empty()) CK(cudaMemcpyAsync(e->d_order, ord.data(), sizeof(int) * ord.size(), cudaMemcpyHostToDevice, e->st));
    CK(cudaStreamSynchronize(e->st));
    return 0;
}

// The fused finish kernel leaves the owned tetrads' new coordinates in crd2 (other CTAs may still
// be reading crd).  The ED pass consumes them from there; every other consumer calls this first.
int normalize(edmd_engine* e) {
    if (!e->latest_in_crd2) return 0;
    const size_t off = (size_t)e->t0 * 3 * e->S, cnt = (size_t)(e->t1 - e->t0) * 3 * e->S;
    if (cnt) CK(cudaMemcpyAsync(e->crd + off, e->crd2 + off, sizeof(double) * cnt, cudaMemcpyDeviceToDevice, e->st));
    e->latest_in_crd2 = false;
    return 0;
}

int do_ed(edmd_engine* e) {
    timer_begin(e, T_ED);
    const double* src = e->latest_in_crd2 ? e->crd2 : e->crd;
    e->latest_in_crd2 = false;
    // The warp-per-tetrad projection stages a parameter set in shared memory once per run of tetrads that
    // share it: worth it when sets are shared (replicated DNA), not when every tetrad has its own.
    // The choice depends on the whole system (n, P, NEV), never on the shard: every rank of a multi-GPU run uses the
    // same kernel, so the results stay identical for every world size.
    const bool shared_sets = (long long)e->n >= 8LL * e->P;
    const bool warp = e->ed_kernel == 2 || (e->ed_kernel == 0 && shared_sets);
    // The tensor-core projection (ed_mma.cu) groups eight tetrads of one parameter set per MMA tile.
    // One warp carries a whole group there (no barrier inside a run), which needs a few thousand groups to fill the
    // machine: small systems keep the warp-per-tetrad kernel.
    const bool mma = ed_project_mma_supported(e->ps) &&
                     (e->ed_kernel == 3 || (e->ed_kernel == 0 && shared_sets && e->ed_auto_mma && e->n >= 4096));
    timer_begin(e, T_SUP);
    CK(launch_superpose(e->ps, e->d_param_id, src, e->sup, e->t0, e->t1 - e->t0, e->sup_variant, e->sms, e->st));
    timer_end(e, T_SUP, 0);
    if (mma)
        CK(launch_ed_project_mma(e->ps, e->d_param_id, src, e->crd, e->fed, e->e_ed, e->sup, e->d_order, e->sms,
                                 e->t1 - e->t0, e->sp.scaled, e->st));
    else
        CK(launch_ed_project(e->ps, e->d_param_id, src, e->crd, e->fed, e->e_ed, e->sup, e->d_order, e->sms, e->t0,
                             e->t1 - e->t0, e->sp.scaled, warp, e->st));
    timer_end(e, T_ED, 2);
    return 0;
}
// which tetrads the culled distance pass needs bounding spheres for: the ends of this engine's pairs
int ensure_need(edmd_engine* e) {
    if (e->need_valid) return 0;
    if (!e->d_need) {
        CK(dev_alloc(&e->d_need, (size_t)e->n)); CK(dev_alloc(&e->d_need_list, (size_t)e->n + 1));
    }
    int* cnt = e->d_need_list + e->n;
    CK(launch_mark_needed(e->d_pairs, e->p0, e->p1, e->d_need, e->d_need_list, cnt, e->n, e->st));
    e->need_count = 0;
    if (e->p1 > e->p0) CK(cudaMemcpyAsync(&e->need_count, cnt, sizeof(int), cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    e->launches += 2;
    e->need_valid = true;
    return 0;
}
int do_scan(edmd_engine* e) {
    const double cut_eff = nb_effective_cut2(e->sp);
    // one memset for the per-step scratch of the NB phase: tetrad marks, marked-list counter, work-item counter,
    // distance-pass item counter (nb.cu: launch_nb_finish_split)
    CK(cudaMemsetAsync(e->d_mark, 0, sizeof(int) * ((size_t)e->n + 4), e->st));
    if (e->scan_mode == 4 && cut_eff >= 1e-3 && cut_eff < 1e6) {
        if (int rc = ensure_need(e)) return rc;
        timer_begin(e, T_SCAN);
        CK(launch_nb_scan_cull(e->crd, e->d_natoms, e->d_pairs, cur_hits(e), e->p0, e->p1, e->n, e->S, cut_eff, e->sms,
                               e->d_gb, e->d_pmask, e->pmask_cap, reinterpret_cast<unsigned*>(e->d_mark + e->n + 2), e->cull_dense_min,
                               e->d_param_id, e->d_perm, e->d_need_list, e->need_count,
                               e->peer_mode ? e->d_hit_tbl : nullptr, e->world, e->st));
        // kernels of the pass: bounding spheres, pair selection, block kernel (sparse masks), tiled kernel (dense masks)
        timer_end(e, T_SCAN, e->p1 > e->p0 ? (e->need_count > 0) + 1 + (e->cull_dense_min > 1) + (e->cull_dense_min <= 64) : 0);
        e->hits_valid = true;
        return 0;
    }
    timer_begin(e, T_SCAN);
    CK(launch_nb_scan(e->crd, e->d_natoms, e->d_pairs, cur_hits(e), e->p0, e->p1, e->S,
                      nb_effective_cut2(e->sp), e->scan_mode, e->sms,
                      e->peer_mode ? e->d_hit_tbl : nullptr, e->world, e->st));
    timer_end(e, T_SCAN, e->p1 > e->p0 ? 1 : 0);
    e->hits_valid = true;
    return 0;
}
int do_accumulate(edmd_engine* e) {
    CK(cudaMemsetAsync(e->d_dirty, 1, (size_t)e->n, e->st));
    timer_begin(e, T_ACC);
    CK(launch_nb_accumulate(e->crd, e->d_natoms, e->d_param_id, e->d_chg, cur_hits(e), e->d_adj_off,
                            e->d_adj_partner, e->d_adj_pair, e->d_perm, e->d_grp, e->fnb, e->e_nb, e->d_epart,
                            reinterpret_cast<unsigned*>(e->d_mark + e->n + 1), e->t0, e->t1 - e->t0, e->S, e->sms, e->sp,
                            e->nb_clip, e->st));
    timer_end(e, T_ACC, 2);
    return 0;
}
// accumulate + clip + velocity + coordinate update in one kernel (edmd_step, edmd_nb_finish)
int do_integrate(edmd_engine* e, int do_vel, int do_crd);
int do_finish(edmd_engine* e, long* bump = nullptr) {
    if (int rc = normalize(e)) return rc;
    if (!e->nb_clip) {   // unclipped forces requested: run the two unfused kernels
        if (int rc = do_accumulate(e)) return rc;
        if (int rc = do_integrate(e, 1, 1)) return rc;
    } else {
        timer_begin(e, T_ACC);
        CK(launch_nb_finish_split(e->crd, e->d_natoms, e->d_param_id, e->d_chg, cur_hits(e), e->d_pairs, e->np, e->d_adj_off,
                                  e->d_adj_partner, e->d_adj_pair, e->d_perm, e->d_grp, e->fnb, e->e_nb, e->d_epart, e->ps,
                                  e->vel, e->crd2, e->fed, e->rnd_zero ? nullptr : e->rnd, e->temp, e->d_mark, e->d_dirty,
                                  e->n, e->t0, e->t1 - e->t0, e->S, e->sms, e->sp, bump, (long)e->n, e->st));
        timer_end(e, T_ACC, e->np > 0 ? 4 : 1);
        e->latest_in_crd2 = true;
    }
    // Peer mode: the masks were stored by every rank's distance kernel; now that this rank has consumed them
    // it clears the buffer.  The peers write into it again only after the next step's first barrier, which
    // this rank joins after the memset (stream order).
    if (e->peer_mode && e->np > 0)
        CK(cudaMemsetAsync(e->d_hitx, 0, sizeof(unsigned long long) * (size_t)e->np, e->st));
    return 0;
}
// per-parameter-set table of Langevin noise amplitudes (rng.cu); depends on gamma, kT and dt
int ensure_noise(edmd_engine* e) {
    if (e->noise_valid) return 0;
    if (!e->d_noise) CK(dev_alloc(&e->d_noise, (size_t)e->P * 3 * e->S));
    CK(launch_noise_table(e->ps, e->P, e->d_noise, e->sp, e->st));
    e->launches++;
    e->noise_valid = true;
    return 0;
}
int do_random(edmd_engine* e, int mode, long time0, int rank, long calls_before, const long* calls_dev = nullptr,
              cudaStream_t on = nullptr) {
    const bool side = on != nullptr;          // the step runs the generator on a branch beside the ED / distance kernels
    if (!on) on = e->st;
    if (mode == 0) {
        if (!e->rnd_zero) {
            CK(cudaMemsetAsync(e->rnd, 0, sizeof(double) * (size_t)e->n * 3 * e->S, e->st));
            e->rnd_zero = true;
        }
        return 0;
    }
    if (mode != 1) return fail(EDMD_ERR_ARG, "random mode %d unknown (0 = off, 1 = reference stream)", mode);
    if (int rc = ensure_noise(e)) return rc;
    if (!side) timer_begin(e, T_RNG);
    // beside other kernels: a resident grid of rng_ctas_per_sm CTAs per SM, so that the generator takes a fixed
    // slice of every SM for most of the step instead of the whole machine for a part of it
    CK(launch_random_terms(e->ps, e->d_noise, e->d_param_id, e->rnd, e->t0, e->t1 - e->t0, e->sp, time0, rank,
                           calls_before + e->t0, calls_dev, side ? e->sms * e->rng_ctas_per_sm : 0, on));
    if (!side) timer_end(e, T_RNG, 1); else e->launches++;
    e->rnd_zero = false;
    return 0;
}
int do_integrate(edmd_engine* e, int do_vel, int do_crd) {
    timer_begin(e, T_INT);
    CK(launch_integrate(e->ps, e->d_param_id, e->vel, e->crd, e->fed, e->rnd_zero ? nullptr : e->rnd, e->fnb,
                        e->temp, e->t0, e->t1 - e->t0, do_vel, do_crd, e->sp, e->st));
    timer_end(e, T_INT, 1);
    return 0;
}

// a snapshot still draining to the host reads the staging buffer: whoever reuses it waits (in stream order)
int snap_fence(edmd_engine* e) {
    for (int k = 0; k < 2; k++)
        if (e->snap_pending[k]) CK(cudaStreamWaitEvent(e->st, e->ev_snap[k], 0));
    return 0;
}

// planes -> interleaved flat on the device (slot of d_stage), then to the pinned slot
int stage_out(edmd_engine* e, const double* planes, int slot) {
    if (int rc = snap_fence(e)) return rc;
    double* d = e->d_stage + (size_t)slot * e->total3;
    CK(launch_unpack_planes(planes, e->d_off3, e->d_natoms, d, e->n, e->S, e->st));
    e->launches++;
    CK(cudaMemcpyAsync(e->h_stage + (size_t)slot * e->total3, d, sizeof(double) * e->total3,
                       cudaMemcpyDeviceToHost, e->st));
    return 0;
}
int stage_in(edmd_engine* e, double* planes, int slot) {
    if (int rc = snap_fence(e)) return rc;
    double* d = e->d_stage + (size_t)slot * e->total3;
    CK(cudaMemcpyAsync(d, e->h_stage + (size_t)slot * e->total3, sizeof(double) * e->total3,
                       cudaMemcpyHostToDevice, e->st));
    CK(launch_pack_planes(d, e->d_off3, e->d_natoms, planes, e->n, e->S, e->st));
    e->launches++;
    return 0;
}

}  // namespace

extern "C" {

const char* edmd_last_error(void) { return g_err.c_str(); }
int edmd_abi_version(void) { return EDMD_B200_ABI_VERSION; }

int edmd_create(const edmd_params* p, int device, edmd_engine** out) {
    if (!out) return fail(EDMD_ERR_ARG, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return fail(EDMD_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                    err != cudaSuccess ? cudaGetErrorString(err) : "device count 0");
    if (device < 0 || device >= count) return fail(EDMD_ERR_ARG, "device %d out of range [0,%d)", device, count);
    CK(cudaSetDevice(device));
    edmd_engine* e = new edmd_engine();
    e->device = device;
    cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, device);
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));     // (numerically lower = more urgent)
    CK(cudaStreamCreateWithPriority(&e->st, cudaStreamNonBlocking, prio_hi));
    CK(cudaStreamCreateWithPriority(&e->st_rng, cudaStreamNonBlocking, prio_lo));
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    // EDMD::EDMD defaults (edmd.cpp:23-43)
    e->sp = SimParams{0.002 * 20.455, 2.0 / 20.455, 0.2 * 20.455, 300.0, 0.002 * 300.0, 30.0, 10.0, 5.0, 100.0, 0.0};
    if (p) {
        e->sp.dt = p->dt; e->sp.gamma = p->gamma; e->sp.tautp = p->tautp; e->sp.temperature = p->temperature;
        e->sp.scaled = p->scaled; e->sp.mole_cutoff = p->mole_Cutoff; e->sp.atom_cutoff = p->atom_Cutoff;
        e->sp.mole_least = p->mole_Least;
    }
    // tuning switches for A/B measurements (none changes a result bit)
    if (const char* v = getenv("EDMD_SUPERPOSE")) { const int k = atoi(v); if (k >= 0 && k <= 2) e->sup_variant = k; }
    if (const char* v = getenv("EDMD_RNG_CTAS")) { const int k = atoi(v); if (k >= 1 && k <= 16) e->rng_ctas_per_sm = k; }
    if (const char* v = getenv("EDMD_RNG_BRANCH")) e->rng_branch = atoi(v) != 0;
    if (const char* v = getenv("EDMD_ED_MMA")) e->ed_auto_mma = atoi(v) != 0;
    if (const char* v = getenv("EDMD_CULL_DENSE")) { const int k = atoi(v); if (k >= 0 && k <= 65) e->cull_dense_min = k; }
    *out = e;
    return 0;
}

void edmd_destroy(edmd_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (int i = 0; i < T_COUNT; i++) timer_fold(e, i);
    free_system(e);
    if (e->span0) { cudaEventDestroy(e->span0); cudaEventDestroy(e->span1); }
    for (cudaEvent_t ev : e->marks) cudaEventDestroy(ev);
    if (e->st_copy) cudaStreamDestroy(e->st_copy);
    if (e->st_rng) { cudaStreamDestroy(e->st_rng); cudaEventDestroy(e->ev_fork); cudaEventDestroy(e->ev_join); }
    if (e->st_out) { cudaStreamDestroy(e->st_out); cudaEventDestroy(e->ev_staged); cudaEventDestroy(e->ev_snap[0]); cudaEventDestroy(e->ev_snap[1]); }
    if (e->ev_crd_in) { cudaEventDestroy(e->ev_crd_in); cudaEventDestroy(e->ev_vel_in); }
    if (e->st) cudaStreamDestroy(e->st);
    delete e;
}

int edmd_set_params(edmd_engine* e, const edmd_params* p) {
    if (!e || !p) return fail(EDMD_ERR_ARG, "null argument");
    e->sp.dt = p->dt; e->sp.gamma = p->gamma; e->sp.tautp = p->tautp; e->sp.temperature = p->temperature;
    e->sp.scaled = p->scaled; e->sp.mole_cutoff = p->mole_Cutoff; e->sp.atom_cutoff = p->atom_Cutoff;
    e->sp.mole_least = p->mole_Least;
    e->hits_valid = false;
    e->noise_valid = false;
    drop_graph(e);
    return 0;
}

int edmd_set_nb_consts(edmd_engine* e, double krep, double qfac) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    e->sp.krep = krep; e->sp.qfac = qfac;
    e->hits_valid = false;
    drop_graph(e);
    return 0;
}

int edmd_set_nb_clip(edmd_engine* e, int on) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    e->nb_clip = on ? 1 : 0;
    drop_graph(e);
    return 0;
}

int edmd_set_ed_kernel(edmd_engine* e, int which) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    if (which < 0 || which > 3) return fail(EDMD_ERR_ARG, "ED kernel %d unknown (0 automatic, 1 thread per atom, 2 warp per tetrad, 3 tensor core)", which);
    e->ed_kernel = which;
    drop_graph(e);
    return 0;
}

int edmd_set_nb_scan_mode(edmd_engine* e, int mode) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    if (mode < 0 || mode > 4 || mode == 3) return fail(EDMD_ERR_ARG, "scan mode %d unknown (0 direct, 1 gram, 2 fp32 prefilter, 4 gram + culling)", mode);
    e->scan_mode = mode;
    drop_graph(e);
    return 0;
}

int edmd_upload_tetrads(edmd_engine* e, const edmd_tetrad* tets, int n, const int* bp_atoms) {
    if (!e || !tets || n <= 0) return fail(EDMD_ERR_ARG, "bad arguments to edmd_upload_tetrads");
    CK(cudaSetDevice(e->device));
    free_system(e);

    int max_atoms = 0, max_nev = 0;
    for (int t = 0; t < n; t++) {
        if (tets[t].num_Atoms <= 0 || tets[t].num_Evecs < 0) return fail(EDMD_ERR_ARG, "tetrad %d: bad sizes", t);
        max_atoms = std::max(max_atoms, tets[t].num_Atoms);
        max_nev = std::max(max_nev, tets[t].num_Evecs);
    }
    const int S = (max_atoms + 31) & ~31;
    if (S > kMaxStride)
        return fail(EDMD_ERR_LIMIT, "tetrad with %d atoms exceeds the engine limit of %d atoms per tetrad",
                    max_atoms, kMaxStride);
    if (max_nev > 64) return fail(EDMD_ERR_LIMIT, "%d eigenvectors exceeds the engine limit of 64", max_nev);
    const int NEV = std::max(max_nev, 1);

    // de-duplicate parameter sets: pointer identity first, then content
    std::unordered_map<PtrKey, int, PtrKeyHash> by_ptr;
    std::unordered_map<uint64_t, std::vector<int> > by_hash;   // hash -> parameter-set ids
    std::vector<int> rep;                                       // representative tetrad of each set
    e->param_id_h.assign(n, 0); e->natoms_h.assign(n, 0); e->nev_h.assign(n, 0);
    e->off3_h.assign(n + 1, 0);
    for (int t = 0; t < n; t++) {
        const edmd_tetrad& T = tets[t];
        e->natoms_h[t] = T.num_Atoms; e->nev_h[t] = T.num_Evecs;
        e->off3_h[t + 1] = e->off3_h[t] + 3LL * T.num_Atoms;
        const double* evec0 = T.num_Evecs ? T.eigenvectors[0] : nullptr;
        PtrKey pk{T.avg, T.masses, T.abq, T.eigenvalues, evec0, T.num_Atoms, T.num_Evecs};
        auto it = by_ptr.find(pk);
        if (it != by_ptr.end()) { e->param_id_h[t] = it->second; continue; }
        const size_t b3 = sizeof(double) * 3 * T.num_Atoms;
        uint64_t h = 0xcbf29ce484222325ULL ^ ((uint64_t)T.num_Atoms << 32) ^ (uint64_t)T.num_Evecs;
        h = hash_bytes(h, T.avg, b3); h = hash_bytes(h, T.masses, b3); h = hash_bytes(h, T.abq, b3);
        h = hash_bytes(h, T.eigenvalues, sizeof(double) * T.num_Evecs);
        if (evec0) h = hash_bytes(h, evec0, b3 * T.num_Evecs);
        int found = -1;
        for (int cand : by_hash[h]) {
            const edmd_tetrad& R = tets[rep[cand]];
            if (R.num_Atoms == T.num_Atoms && R.num_Evecs == T.num_Evecs && !memcmp(R.avg, T.avg, b3) &&
                !memcmp(R.masses, T.masses, b3) && !memcmp(R.abq, T.abq, b3) &&
                !memcmp(R.eigenvalues, T.eigenvalues, sizeof(double) * T.num_Evecs) &&
                (!evec0 || !memcmp(R.eigenvectors[0], evec0, b3 * T.num_Evecs))) { found = cand; break; }
        }
        if (found < 0) { found = (int)rep.size(); rep.push_back(t); by_hash[h].push_back(found); }
        by_ptr[pk] = found;
        e->param_id_h[t] = found;
    }
    const int P = (int)rep.size();
    e->n = n; e->S = S; e->P = P; e->NEV = NEV; e->total3 = e->off3_h[n];

    // parameter sets -> plane layout
    std::vector<int> ps_na(P), ps_nev(P);
    std::vector<double> avg((size_t)P * 3 * S, 0.0), mass((size_t)P * 3 * S, 1.0), chg((size_t)P * S, 0.0);
    std::vector<double> eval((size_t)P * NEV, 1.0), evec((size_t)P * NEV * 3 * S, 0.0);
    for (int p = 0; p < P; p++) {
        const edmd_tetrad& T = tets[rep[p]];
        ps_na[p] = T.num_Atoms; ps_nev[p] = T.num_Evecs;
        for (int a = 0; a < T.num_Atoms; a++) {
            for (int c = 0; c < 3; c++) {
                avg[((size_t)p * 3 + c) * S + a] = T.avg[3 * a + c];
                mass[((size_t)p * 3 + c) * S + a] = T.masses[3 * a + c];
            }
            chg[(size_t)p * S + a] = T.abq[3 * a + 2];
        }
        for (int j = 0; j < T.num_Evecs; j++) {
            eval[(size_t)p * NEV + j] = T.eigenvalues[j];
            const double* row = T.eigenvectors[0] + (size_t)j * 3 * T.num_Atoms;
            for (int a = 0; a < T.num_Atoms; a++)
                for (int c = 0; c < 3; c++) evec[(((size_t)p * NEV + j) * 3 + c) * S + a] = row[3 * a + c];
        }
    }

    // slot permutation for the culling scan and the force pass: recursive coordinate bisection of each
    // average structure into 8 groups of <= 32 atoms (slots 32g..32g+31 of kMaxStride = 256 slots, whatever
    // the atom stride S is), padded with -1
    std::vector<int> perm((size_t)P * kMaxStride, -1);
    for (int p = 0; p < P; p++) {
        const edmd_tetrad& T = tets[rep[p]];
        std::vector<int> idx(T.num_Atoms);
        for (int a = 0; a < T.num_Atoms; a++) idx[a] = a;
        // split [lo,hi) into `parts` groups along the longest axis, sizes as even as possible
        struct Job { int lo, hi, parts, first; };
        std::vector<Job> jobs{{0, T.num_Atoms, 8, 0}};
        while (!jobs.empty()) {
            Job j = jobs.back(); jobs.pop_back();
            if (j.parts == 1) {
                // slot 0 of the group = its 1-centre (smallest maximum distance to the other members):
                // group_bounds_kernel centres the group's bounding sphere on that atom
                int best = j.lo; double best_r = 1e300;
                for (int q = j.lo; q < j.hi; q++) {
                    double r2 = 0.0;
                    for (int u = j.lo; u < j.hi; u++) {
                        double d2 = 0.0;
                        for (int c = 0; c < 3; c++) { const double d = T.avg[3 * idx[q] + c] - T.avg[3 * idx[u] + c]; d2 += d * d; }
                        r2 = std::max(r2, d2);
                    }
                    if (r2 < best_r) { best_r = r2; best = q; }
                }
                if (j.hi > j.lo) std::swap(idx[j.lo], idx[best]);
                for (int q = j.lo; q < j.hi && q - j.lo < 32; q++) perm[(size_t)p * kMaxStride + 32 * j.first + (q - j.lo)] = idx[q];
                continue;
            }
            double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
            for (int q = j.lo; q < j.hi; q++)
                for (int c = 0; c < 3; c++) { const double v = T.avg[3 * idx[q] + c]; mn[c] = std::min(mn[c], v); mx[c] = std::max(mx[c], v); }
            int ax = 0;
            for (int c = 1; c < 3; c++) if (mx[c] - mn[c] > mx[ax] - mn[ax]) ax = c;
            std::sort(idx.begin() + j.lo, idx.begin() + j.hi, [&](int a, int b) { return T.avg[3 * a + ax] < T.avg[3 * b + ax]; });
            const int half = j.parts / 2;
            int cnt = j.hi - j.lo;
            int left = (cnt + 1) / 2;
            if (left > 32 * half) left = 32 * half;
            if (cnt - left > 32 * half) left = cnt - 32 * half;
            jobs.push_back({j.lo, j.lo + left, half, j.first});
            jobs.push_back({j.lo + left, j.hi, half, j.first + half});
        }
    }
    // inverse table for the force pass: slot group of every atom, eight atoms (32c + l, c = 0..7) per word
    std::vector<unsigned long long> grp((size_t)P * 32, ~0ull);
    for (int p = 0; p < P; p++)
        for (int slot = 0; slot < kMaxStride; slot++) {
            const int a = perm[(size_t)p * kMaxStride + slot];
            if (a < 0) continue;
            unsigned long long& wd = grp[(size_t)p * 32 + (a & 31)];
            const int sh = 8 * (a >> 5);
            wd = (wd & ~(0xffull << sh)) | ((unsigned long long)(slot >> 5) << sh);
        }
    CK(dev_alloc(&e->d_perm, perm.size()));
    CK(cudaMemcpyAsync(e->d_perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice, e->st));
    CK(dev_alloc(&e->d_grp, grp.size()));
    CK(cudaMemcpyAsync(e->d_grp, grp.data(), sizeof(unsigned long long) * grp.size(), cudaMemcpyHostToDevice, e->st));
    CK(dev_alloc(&e->d_natoms, n)); CK(dev_alloc(&e->d_param_id, n)); CK(dev_alloc(&e->d_off3, n + 1));
    CK(dev_alloc(&e->d_ps_na, P)); CK(dev_alloc(&e->d_ps_nev, P));
    CK(dev_alloc(&e->d_avg, avg.size())); CK(dev_alloc(&e->d_mass, mass.size())); CK(dev_alloc(&e->d_chg, chg.size()));
    CK(dev_alloc(&e->d_eval, eval.size())); CK(dev_alloc(&e->d_evec, evec.size()));
    const size_t planes = (size_t)n * 3 * S;
    CK(dev_alloc(&e->crd, planes)); CK(dev_alloc(&e->crd2, planes)); CK(dev_alloc(&e->vel, planes)); CK(dev_alloc(&e->fed, planes));
    CK(dev_alloc(&e->rnd, planes)); CK(dev_alloc(&e->fnb, planes));
    CK(dev_alloc(&e->obs, 4 * (size_t)n));
    e->e_ed = e->obs; e->e_nb = e->obs + n; e->temp = e->obs + 3 * (size_t)n;
    CK(dev_alloc(&e->cen, 4 * (size_t)n));
    CK(dev_alloc(&e->d_mark, 2 * (size_t)n + 8)); CK(dev_alloc(&e->d_dirty, (size_t)n));
    CK(cudaMemsetAsync(e->d_mark, 0, sizeof(int) * (2 * (size_t)n + 8), e->st));
    CK(dev_alloc(&e->d_epart, 16 * (size_t)n));
    CK(cudaMemsetAsync(e->d_epart, 0, sizeof(double) * 16 * (size_t)n, e->st));
    CK(dev_alloc(&e->d_gb, (size_t)n * 36));
    CK(cudaMemsetAsync(e->d_dirty, 1, (size_t)n, e->st));   // unknown force contents: treat as dirty
    CK(dev_alloc(&e->sup, (size_t)kSupWordsHost * n));
    CK(dev_alloc(&e->d_stage, 3 * (size_t)e->total3));
    CK(cudaMallocHost((void**)&e->h_stage, sizeof(double) * 3 * (size_t)e->total3));
    CK(dev_alloc(&e->d_row_count, n)); CK(dev_alloc(&e->d_row_off, n + 1)); CK(dev_alloc(&e->d_adj_off, n + 1));

    CK(cudaMemcpyAsync(e->d_natoms, e->natoms_h.data(), sizeof(int) * n, cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_param_id, e->param_id_h.data(), sizeof(int) * n, cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_off3, e->off3_h.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_ps_na, ps_na.data(), sizeof(int) * P, cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_ps_nev, ps_nev.data(), sizeof(int) * P, cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_avg, avg.data(), sizeof(double) * avg.size(), cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_mass, mass.data(), sizeof(double) * mass.size(), cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_chg, chg.data(), sizeof(double) * chg.size(), cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_eval, eval.data(), sizeof(double) * eval.size(), cudaMemcpyHostToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_evec, evec.data(), sizeof(double) * evec.size(), cudaMemcpyHostToDevice, e->st));
    for (double* buf : {e->vel, e->fed, e->rnd, e->fnb, e->crd2})     // (padding slots beyond a tetrad's atoms stay zero)
        CK(cudaMemsetAsync(buf, 0, sizeof(double) * planes, e->st));
    CK(cudaMemsetAsync(e->obs, 0, sizeof(double) * 4 * n, e->st));
    CK(cudaStreamSynchronize(e->st));   // host vectors go out of scope
    e->rnd_zero = true;

    e->ps = ParamSets{e->d_ps_na, e->d_ps_nev, e->d_avg, e->d_mass, e->d_chg, e->d_eval, e->d_evec, S, NEV};
    CK(dev_alloc(&e->d_pspre, (size_t)P * 4));
    e->ps.pre = e->d_pspre;
    CK(launch_ps_precompute(e->ps, P, e->d_pspre, e->st));
    e->t0 = 0; e->t1 = n; e->p0 = e->p1 = 0; e->shard_pairs_custom = false;
    if (int rc = upload_order(e)) return rc;

    if (bp_atoms) {
        std::vector<int> off(n + 4, 0);
        for (int b = 0; b < n + 3; b++) off[b + 1] = off[b] + bp_atoms[b];
        for (int t = 0; t < n; t++)
            if (off[t + 4] - off[t] != e->natoms_h[t])
                return fail(EDMD_ERR_ARG, "bp_atoms does not match tetrad %d: %d atoms vs %d (io.cpp:267-269)",
                            t, off[t + 4] - off[t], e->natoms_h[t]);
        CK(dev_alloc(&e->d_bp_off, n + 4));
        CK(cudaMemcpy(e->d_bp_off, off.data(), sizeof(int) * (n + 4), cudaMemcpyHostToDevice));
        e->have_bp = true;
    }
    return edmd_set_state(e, tets, n);
}

// True when field `which` (0 coordinates, 1 velocities) of all tetrads forms one dense block in
// tetrad order (a caller that allocates its Tetrad arrays from one slab — the copy can then
// skip the gather through the engine's staging buffer).
static bool dense_field(const edmd_engine* e, const edmd_tetrad* tets, int which) {
    const double* base = which ? tets[0].velocities : tets[0].coordinates;
    for (int t = 1; t < e->n; t++) {
        const double* p = which ? tets[t].velocities : tets[t].coordinates;
        if (p != base + e->off3_h[t]) return false;
    }
    return true;
}

int edmd_set_state(edmd_engine* e, const edmd_tetrad* tets, int n) {
    if (int rc = check_ready(e, false)) return rc;
    if (!tets || n != e->n) return fail(EDMD_ERR_ARG, "edmd_set_state: n=%d but engine holds %d tetrads", n, e->n);
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    if (int rc = snap_fence(e)) return rc;
    timer_begin(e, T_LAYOUT);
    for (int which = 0; which < 2; which++) {
        double* planes = which ? e->vel : e->crd;
        if (dense_field(e, tets, which)) {       // straight from the caller's memory (fast if pinned)
            double* d = e->d_stage + (size_t)which * e->total3;
            CK(cudaMemcpyAsync(d, which ? tets[0].velocities : tets[0].coordinates, sizeof(double) * e->total3,
                               cudaMemcpyHostToDevice, e->st));
            CK(launch_pack_planes(d, e->d_off3, e->d_natoms, planes, e->n, e->S, e->st));
            e->launches++;
        } else {
            for (int t = 0; t < n; t++)
                memcpy(e->h_stage + (size_t)which * e->total3 + e->off3_h[t],
                       which ? tets[t].velocities : tets[t].coordinates, sizeof(double) * 3 * e->natoms_h[t]);
            if (int rc = stage_in(e, planes, which)) return rc;
        }
    }
    timer_end(e, T_LAYOUT, 0);
    CK(cudaStreamSynchronize(e->st));
    e->hits_valid = false;
    return 0;
}

int edmd_get_state(edmd_engine* e, edmd_tetrad* tets, int n) {
    if (int rc = check_ready(e, false)) return rc;
    if (!tets || n != e->n) return fail(EDMD_ERR_ARG, "edmd_get_state: n mismatch");
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    if (int rc = snap_fence(e)) return rc;
    bool dense[2];
    for (int which = 0; which < 2; which++) {
        dense[which] = dense_field(e, tets, which);
        const double* planes = which ? e->vel : e->crd;
        if (dense[which]) {
            double* d = e->d_stage + (size_t)which * e->total3;
            CK(launch_unpack_planes(planes, e->d_off3, e->d_natoms, d, e->n, e->S, e->st));
            e->launches++;
            CK(cudaMemcpyAsync(which ? tets[0].velocities : tets[0].coordinates, d, sizeof(double) * e->total3,
                               cudaMemcpyDeviceToHost, e->st));
        } else if (int rc = stage_out(e, planes, which)) return rc;
    }
    CK(cudaStreamSynchronize(e->st));
    for (int which = 0; which < 2; which++) {
        if (dense[which]) continue;
        for (int t = 0; t < n; t++)
            memcpy(which ? tets[t].velocities : tets[t].coordinates,
                   e->h_stage + (size_t)which * e->total3 + e->off3_h[t], sizeof(double) * 3 * e->natoms_h[t]);
    }
    return 0;
}

// The owned tetrads [t0,t1) only: what one rank of a multi-process run moves per step when every rank keeps
// its own block of the host Tetrad array current (the reference's workers received everything from the
// master, master.cpp:279; with owner-computes each rank needs only its block).  Dense host fields required.
int edmd_set_state_owned(edmd_engine* e, const edmd_tetrad* tets, int n) {
    if (int rc = check_ready(e, false)) return rc;
    if (!tets || n != e->n) return fail(EDMD_ERR_ARG, "edmd_set_state_owned: n mismatch");
    if (!dense_field(e, tets, 0) || !dense_field(e, tets, 1))
        return fail(EDMD_ERR_ARG, "edmd_set_state_owned needs the coordinates (velocities) of all tetrads in one block each");
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    if (int rc = snap_fence(e)) return rc;
    const int cnt = e->t1 - e->t0;
    if (cnt <= 0) return 0;
    const long long o0 = e->off3_h[e->t0], o1 = e->off3_h[e->t1];
    for (int which = 0; which < 2; which++) {
        double* d = e->d_stage + (size_t)which * e->total3;
        const double* h = which ? tets[0].velocities : tets[0].coordinates;
        CK(cudaMemcpyAsync(d + o0, h + o0, sizeof(double) * (o1 - o0), cudaMemcpyHostToDevice, e->st));
        CK(launch_pack_planes(d, e->d_off3 + e->t0, e->d_natoms + e->t0, (which ? e->vel : e->crd) + (size_t)e->t0 * 3 * e->S,
                              cnt, e->S, e->st));
        e->launches++;
    }
    e->hits_valid = false;
    return 0;
}

int edmd_get_state_owned(edmd_engine* e, edmd_tetrad* tets, int n) {
    if (int rc = check_ready(e, false)) return rc;
    if (!tets || n != e->n) return fail(EDMD_ERR_ARG, "edmd_get_state_owned: n mismatch");
    if (!dense_field(e, tets, 0) || !dense_field(e, tets, 1))
        return fail(EDMD_ERR_ARG, "edmd_get_state_owned needs the coordinates (velocities) of all tetrads in one block each");
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    if (int rc = snap_fence(e)) return rc;
    const int cnt = e->t1 - e->t0;
    if (cnt > 0) {
        const long long o0 = e->off3_h[e->t0], o1 = e->off3_h[e->t1];
        for (int which = 0; which < 2; which++) {
            double* d = e->d_stage + (size_t)which * e->total3;
            double* h = which ? tets[0].velocities : tets[0].coordinates;
            CK(launch_unpack_planes((which ? e->vel : e->crd) + (size_t)e->t0 * 3 * e->S, e->d_off3 + e->t0, e->d_natoms + e->t0,
                                    d, cnt, e->S, e->st));
            e->launches++;
            CK(cudaMemcpyAsync(h + o0, d + o0, sizeof(double) * (o1 - o0), cudaMemcpyDeviceToHost, e->st));
        }
    }
    CK(cudaStreamSynchronize(e->st));
    return peer_check(e);
}

int edmd_set_state_flat(edmd_engine* e, const double* crd, const double* vel) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    if (crd) { memcpy(e->h_stage, crd, sizeof(double) * e->total3); if (int rc = stage_in(e, e->crd, 0)) return rc; }
    if (vel) { memcpy(e->h_stage + e->total3, vel, sizeof(double) * e->total3); if (int rc = stage_in(e, e->vel, 1)) return rc; }
    CK(cudaStreamSynchronize(e->st));
    e->hits_valid = false;
    return 0;
}

int edmd_get_state_flat(edmd_engine* e, double* crd, double* vel) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    if (crd) if (int rc = stage_out(e, e->crd, 0)) return rc;
    if (vel) if (int rc = stage_out(e, e->vel, 1)) return rc;
    CK(cudaStreamSynchronize(e->st));
    if (crd) memcpy(crd, e->h_stage, sizeof(double) * e->total3);
    if (vel) memcpy(vel, e->h_stage + e->total3, sizeof(double) * e->total3);
    return 0;
}

int edmd_get_forces_flat(edmd_engine* e, double* ed, double* ed_energy, double* rnd, double* nb,
                         double* nb_energy, double* temperature) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    if (ed) if (int rc = stage_out(e, e->fed, 0)) return rc;
    if (rnd) if (int rc = stage_out(e, e->rnd, 1)) return rc;
    if (nb) if (int rc = stage_out(e, e->fnb, 2)) return rc;
    if (ed_energy) CK(cudaMemcpyAsync(ed_energy, e->e_ed, sizeof(double) * e->n, cudaMemcpyDeviceToHost, e->st));
    if (nb_energy) CK(cudaMemcpyAsync(nb_energy, e->e_nb, sizeof(double) * 2 * e->n, cudaMemcpyDeviceToHost, e->st));
    if (temperature) CK(cudaMemcpyAsync(temperature, e->temp, sizeof(double) * e->n, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    if (ed) memcpy(ed, e->h_stage, sizeof(double) * e->total3);
    if (rnd) memcpy(rnd, e->h_stage + e->total3, sizeof(double) * e->total3);
    if (nb) memcpy(nb, e->h_stage + 2 * e->total3, sizeof(double) * e->total3);
    return 0;
}

int edmd_get_forces(edmd_engine* e, edmd_tetrad* tets, int n) {
    if (int rc = check_ready(e, false)) return rc;
    if (!tets || n != e->n) return fail(EDMD_ERR_ARG, "edmd_get_forces: n mismatch");
    std::vector<double> eed(n), enb(2 * (size_t)n), tmp(n);
    std::vector<double> ed(e->total3), rn(e->total3), nb(e->total3);
    if (int rc = edmd_get_forces_flat(e, ed.data(), eed.data(), rn.data(), nb.data(), enb.data(), tmp.data())) return rc;
    for (int t = 0; t < n; t++) {
        const int n3 = 3 * e->natoms_h[t];
        const long long o = e->off3_h[t];
        memcpy(tets[t].ED_Forces, ed.data() + o, sizeof(double) * n3);
        tets[t].ED_Forces[n3] = eed[t];
        memcpy(tets[t].random_Terms, rn.data() + o, sizeof(double) * n3);
        memcpy(tets[t].NB_Forces, nb.data() + o, sizeof(double) * n3);
        tets[t].NB_Forces[n3] = enb[2 * (size_t)t];
        tets[t].NB_Forces[n3 + 1] = enb[2 * (size_t)t + 1];
        tets[t].temperature = tmp[t];
    }
    return 0;
}

int edmd_set_forces(edmd_engine* e, const edmd_tetrad* tets, int n) {
    if (int rc = check_ready(e, false)) return rc;
    if (!tets || n != e->n) return fail(EDMD_ERR_ARG, "edmd_set_forces: n mismatch");
    CK(cudaSetDevice(e->device));
    for (int t = 0; t < n; t++) {
        const size_t b = sizeof(double) * 3 * e->natoms_h[t];
        memcpy(e->h_stage + e->off3_h[t], tets[t].ED_Forces, b);
        memcpy(e->h_stage + e->total3 + e->off3_h[t], tets[t].random_Terms, b);
        memcpy(e->h_stage + 2 * e->total3 + e->off3_h[t], tets[t].NB_Forces, b);
    }
    if (int rc = stage_in(e, e->fed, 0)) return rc;
    if (int rc = stage_in(e, e->rnd, 1)) return rc;
    if (int rc = stage_in(e, e->fnb, 2)) return rc;
    CK(cudaMemsetAsync(e->d_dirty, 1, (size_t)e->n, e->st));
    CK(cudaStreamSynchronize(e->st));
    e->rnd_zero = false;
    return 0;
}

int edmd_set_pair_builder(edmd_engine* e, int which) {
    if (!e || which < 0 || which > 2) return fail(EDMD_ERR_ARG, "pair builder must be 0 (auto), 1 (sweep) or 2 (cells)");
    e->pair_builder = which;
    return 0;
}

int edmd_build_pairlist(edmd_engine* e, long long* num_pairs) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    const int n = e->n;
    const double cut2 = e->sp.mole_cutoff * e->sp.mole_cutoff;
    timer_begin(e, T_PAIRS);
    CK(launch_centroids(e->crd, e->d_natoms, e->cen, n, e->S, e->st));

    // builder choice: the cell list pays off when the cutoff is small against the extent.  Only the bounding
    // box (7 doubles) and the pair count cross PCIe; sums, offsets and the adjacency stay on the device.
    bool cells = false;
    CellGridHost grid{};
    if (e->pair_builder != 1 && e->sp.mole_cutoff > 0.0 && e->sp.mole_cutoff < 1e8) {
        if (!e->d_bbox) CK(dev_alloc(&e->d_bbox, 8));
        CK(launch_centroid_bbox(e->cen, n, e->d_bbox, e->st));
        double bb[7];
        CK(cudaMemcpyAsync(bb, e->d_bbox, sizeof(bb), cudaMemcpyDeviceToHost, e->st));
        CK(cudaStreamSynchronize(e->st));
        const double* lo = bb; const double* hi = bb + 3;
        const bool finite = bb[6] != 0.0;
        const double h = e->sp.mole_cutoff;
        const double ncell = finite ? (floor((hi[0] - lo[0]) / h) + 1) * (floor((hi[1] - lo[1]) / h) + 1) *
                                          (floor((hi[2] - lo[2]) / h) + 1) : 0.0;
        cells = finite && ncell < 9e18 && (e->pair_builder == 2 || (n >= 4096 && ncell >= 64.0));
        if (cells) {
            if (!e->d_cell) {
                CK(dev_alloc(&e->d_cell, (size_t)n)); CK(dev_alloc(&e->d_cell_sorted, (size_t)n));
                CK(dev_alloc(&e->d_tet, (size_t)n)); CK(dev_alloc(&e->d_tet_sorted, (size_t)n));
            }
            CK(cell_sort_tetrads(e->cen, n, lo, hi, h, &grid, e->d_cell, e->d_cell_sorted, e->d_tet, e->d_tet_sorted,
                                 &e->d_cub_tmp, &e->cub_tmp_bytes, e->st));
        }
    }

    if (cells) CK(launch_cell_rows(e->cen, n, cut2, e->sp.mole_least, grid, e->d_cell_sorted, e->d_tet_sorted,
                                   e->d_row_count, nullptr, nullptr, 0, e->st));
    else CK(launch_pair_rows(e->cen, n, cut2, e->sp.mole_least, e->d_row_count, nullptr, nullptr, 0, e->st));
    CK(scan_row_counts(e->d_row_count, e->d_row_off, n, &e->d_cub_tmp, &e->cub_tmp_bytes, e->st));
    long long np = 0;
    CK(cudaMemcpyAsync(&np, e->d_row_off + n, sizeof(long long), cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    if (np >= (1LL << 31)) return fail(EDMD_ERR_LIMIT, "%lld pairs exceeds the 2^31 pair-index limit", np);
    if (int rc = ensure_pair_capacity(e, np)) return rc;
    if (cells) {
        if (np > e->partners_cap) {
            cudaFree(e->d_partners); cudaFree(e->d_partners_sorted);
            e->d_partners = e->d_partners_sorted = nullptr; e->partners_cap = 0;
            const long long cap = np + np / 4 + 1024;
            CK(dev_alloc(&e->d_partners, (size_t)cap)); CK(dev_alloc(&e->d_partners_sorted, (size_t)cap));
            e->partners_cap = cap;
        }
        CK(launch_cell_rows(e->cen, n, cut2, e->sp.mole_least, grid, e->d_cell_sorted, e->d_tet_sorted,
                            e->d_row_count, e->d_row_off, e->d_partners, 1, e->st));
        CK(sort_rows_and_emit(e->d_partners, e->d_partners_sorted, e->d_row_off, n, np, e->d_pairs,
                              &e->d_cub_tmp, &e->cub_tmp_bytes, e->st));
    } else {
        CK(launch_pair_rows(e->cen, n, cut2, e->sp.mole_least, e->d_row_count, e->d_row_off, e->d_pairs, 1, e->st));
    }
    e->np = np;
    if (int rc = build_adjacency(e)) return rc;
    timer_end(e, T_PAIRS, cells ? 9 : 5);
    e->have_pairs = true; e->hits_valid = false; drop_graph(e);
    if (!e->shard_pairs_custom) { e->p0 = 0; e->p1 = np; }
    else { e->p0 = std::min(e->p0, np); e->p1 = std::min(e->p1, np); }
    if (e->peer_mode && np > 0) CK(cudaMemsetAsync(e->d_hitx, 0, sizeof(unsigned long long) * (size_t)np, e->st));
    if (int rc = rebuild_halo(e)) return rc;
    if (num_pairs) *num_pairs = np;
    return 0;
}

int edmd_get_pairlist(edmd_engine* e, int* pairs, long long cap_pairs) {
    if (int rc = check_ready(e, true)) return rc;
    if (!pairs || cap_pairs < e->np) return fail(EDMD_ERR_ARG, "pair buffer too small: %lld < %lld", cap_pairs, e->np);
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->st));
    if (e->np) CK(cudaMemcpy(pairs, e->d_pairs, sizeof(int2) * e->np, cudaMemcpyDeviceToHost));
    return 0;
}

int edmd_set_pairlist(edmd_engine* e, const int* pairs, long long np) {
    if (int rc = check_ready(e, false)) return rc;
    if (np < 0 || (np > 0 && !pairs)) return fail(EDMD_ERR_ARG, "bad pair list");
    if (np >= (1LL << 31)) return fail(EDMD_ERR_LIMIT, "%lld pairs exceeds the 2^31 pair-index limit", np);
    CK(cudaSetDevice(e->device));
    std::vector<int2> pv((size_t)np);
    for (long long p = 0; p < np; p++) {
        const int a = pairs[2 * p], b = pairs[2 * p + 1];
        if (a < 0 || b < 0 || a >= e->n || b >= e->n || a == b)
            return fail(EDMD_ERR_ARG, "pair %lld = (%d,%d) out of range", p, a, b);
        pv[p] = make_int2(a, b);
    }
    CK(cudaStreamSynchronize(e->st));
    if (int rc = ensure_pair_capacity(e, np)) return rc;
    if (np) CK(cudaMemcpy(e->d_pairs, pv.data(), sizeof(int2) * np, cudaMemcpyHostToDevice));
    e->np = np;
    if (int rc = build_adjacency(e)) return rc;
    CK(cudaStreamSynchronize(e->st));
    e->have_pairs = true; e->hits_valid = false; drop_graph(e);
    if (!e->shard_pairs_custom) { e->p0 = 0; e->p1 = np; }
    else { e->p0 = std::min(e->p0, np); e->p1 = std::min(e->p1, np); }
    if (e->peer_mode && np > 0) CK(cudaMemsetAsync(e->d_hitx, 0, sizeof(unsigned long long) * (size_t)np, e->st));
    return rebuild_halo(e);
}

int edmd_set_shard(edmd_engine* e, int t0, int t1, long long p0, long long p1) {
    if (int rc = check_ready(e, false)) return rc;
    if (t0 < 0 || t1 > e->n || t0 > t1) return fail(EDMD_ERR_ARG, "tetrad shard [%d,%d) out of range", t0, t1);
    if (p0 < 0 || p0 > p1 || (e->have_pairs && p1 > e->np)) return fail(EDMD_ERR_ARG, "pair shard out of range");
    CK(cudaSetDevice(e->device));
    if (e->shard_pairs_custom && t0 == e->t0 && t1 == e->t1 && p0 == e->p0 && p1 == e->p1) return 0;
    if (int rc = normalize(e)) return rc;
    const bool range_changed = (t0 != e->t0 || t1 != e->t1);
    drop_graph(e);
    e->t0 = t0; e->t1 = t1; e->p0 = p0; e->p1 = p1; e->shard_pairs_custom = true;
    e->hits_valid = false;
    e->peer_replicated = (t0 == 0 && t1 == e->n);
    if (range_changed) if (int rc = upload_order(e)) return rc;
    return rebuild_halo(e);
}

int edmd_ed_forces(edmd_engine* e) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    e->hits_valid = false;
    return do_ed(e);
}

int edmd_random_terms(edmd_engine* e, int mode, long time0, int rank, long calls_before) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    return do_random(e, mode, time0, rank, calls_before);
}

int edmd_nb_scan(edmd_engine* e) {
    if (int rc = check_ready(e, true)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    return do_scan(e);
}

int edmd_nb_accumulate(edmd_engine* e) {
    if (int rc = check_ready(e, true)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    return do_accumulate(e);
}

int edmd_nb_forces(edmd_engine* e) {
    if (int rc = check_ready(e, true)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    if (int rc = do_scan(e)) return rc;
    return do_accumulate(e);
}

int edmd_update_velocities(edmd_engine* e) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    return do_integrate(e, 1, 0);
}

int edmd_update_coordinates(edmd_engine* e) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    e->hits_valid = false;
    return do_integrate(e, 0, 1);
}

int edmd_nb_finish(edmd_engine* e) {
    if (int rc = check_ready(e, true)) return rc;
    CK(cudaSetDevice(e->device));
    e->hits_valid = false;
    return do_finish(e);
}

int edmd_integrate(edmd_engine* e) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    e->hits_valid = false;
    return do_integrate(e, 1, 1);
}

int edmd_merge(edmd_engine* e) {
    if (int rc = check_ready(e, false)) return rc;
    if (!e->have_bp) return fail(EDMD_ERR_STATE, "edmd_merge needs bp_atoms (pass it to edmd_upload_tetrads)");
    if (e->n < 4) return fail(EDMD_ERR_LIMIT, "merge needs at least 4 tetrads");
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    timer_begin(e, T_MERGE);
    CK(launch_merge(e->vel, e->crd, e->d_bp_off, e->n, e->S, e->st));
    timer_end(e, T_MERGE, 1);
    e->hits_valid = false;
    return 0;
}

int edmd_get_energies(edmd_engine* e, double out[4]) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    const int n = e->n;
    std::vector<double> eed(n), enb(2 * (size_t)n), tmp(n);
    CK(cudaMemcpyAsync(eed.data(), e->e_ed, sizeof(double) * n, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(enb.data(), e->e_nb, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(tmp.data(), e->temp, sizeof(double) * n, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    // master.cpp:393-401: sequential sums in tetrad order (4 doubles per tetrad; host side like
    // the reference's master, output-frequency only)
    out[0] = out[1] = out[2] = out[3] = 0.0;
    for (int t = 0; t < n; t++) { out[0] += eed[t]; out[1] += enb[2 * (size_t)t]; out[2] += enb[2 * (size_t)t + 1]; out[3] += tmp[t]; }
    out[3] /= n;
    return 0;
}

int edmd_set_rng(edmd_engine* e, long time0, int rank, long calls_before) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    e->time0 = time0; e->rank = rank; e->calls = calls_before;
    drop_graph(e);
    return 0;
}

// One step on this rank.  With peers attached the two exchange points of the path are (1) after the ED pass,
// when every owner's re-embedded coordinates exist: barrier, then pull the non-owned tetrads this rank reads;
// (2) after the distance pass, whose kernels stored their masks into every rank's buffer: barrier, then the
// owners' force pass.  Both are kernels on the engine stream, so the whole sharded step is graph-captured.
//
// The random terms depend on nothing but the call counter and are consumed only by the final update, so the
// generator CAN run on a second stream (a parallel branch of the captured graph, edmd_set_rng_branch) and join
// before the finish kernel.  Measured at 1e5 tetrads: 3.22-3.57 ms per step with the branch (any resident grid
// size) against 3.16 ms back to back — every kernel of the step is bound by latency or issue and wants all the
// warp slots of an SM, so co-running only divides them.  The branch is therefore off by default.
static int one_step(edmd_engine* e, int random_mode, const long* calls_dev) {
    const bool fork = random_mode == 1 && !e->timing && e->rng_branch;
    if (fork) {
        if (int rc = ensure_noise(e)) return rc;      // (its table is built on the main stream, before the fork)
        CK(cudaEventRecord(e->ev_fork, e->st));       // after the previous step's finish kernel, which read e->rnd
        CK(cudaStreamWaitEvent(e->st_rng, e->ev_fork, 0));
        if (int rc = do_random(e, random_mode, e->time0, e->rank, e->calls, calls_dev, e->st_rng)) return rc;
        if (calls_dev) { CK(launch_bump_counter(e->d_calls, (long)e->n, e->st_rng)); e->launches++; }
        CK(cudaEventRecord(e->ev_join, e->st_rng));
    }
    if (int rc = do_ed(e)) return rc;
    if (!fork) {
        if (int rc = do_random(e, random_mode, e->time0, e->rank, e->calls, calls_dev)) return rc;
        if (random_mode == 1 && calls_dev && !e->nb_clip) { CK(launch_bump_counter(e->d_calls, (long)e->n, e->st)); e->launches++; }
    }
    if (e->peer_mode && !e->peer_replicated) {
        if (int rc = peer_barrier(e)) return rc;
        CK(launch_halo_pull(e->crd, e->d_crd_tbl, e->d_halo, e->halo_count, (e->n + e->world - 1) / e->world, e->S,
                            e->sms, e->st));
        if (e->halo_count) e->launches++;
    }
    if (int rc = do_scan(e)) return rc;
    if (e->peer_mode) if (int rc = peer_barrier(e)) return rc;
    if (fork) CK(cudaStreamWaitEvent(e->st, e->ev_join, 0));
    // (the replayed graph's call counter advances inside the finish kernel, after the generator has read it)
    return do_finish(e, (!fork && random_mode == 1 && calls_dev && e->nb_clip) ? e->d_calls : nullptr);
}

int edmd_set_rng_branch(edmd_engine* e, int on) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    if (on < 0 || on > 16) return fail(EDMD_ERR_ARG, "rng branch: 0 = off, 1 = on, 2..16 = on with that many generator CTAs per SM");
    e->rng_branch = on != 0;
    if (on >= 2) e->rng_ctas_per_sm = on;
    drop_graph(e);
    return 0;
}

int edmd_set_graph(edmd_engine* e, int on) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    e->use_graph = on != 0;
    if (!on) drop_graph(e);
    return 0;
}

int edmd_step(edmd_engine* e, int nsteps, int random_mode) {
    if (int rc = check_ready(e, true)) return rc;
    if (random_mode != 0 && random_mode != 1) return fail(EDMD_ERR_ARG, "random mode %d unknown", random_mode);
    CK(cudaSetDevice(e->device));
    const bool graph_ok = e->use_graph && !e->timing && e->nb_clip && nsteps >= 1;
    if (!graph_ok) {
        for (int it = 0; it < nsteps; it++) {
            if (int rc = one_step(e, random_mode, nullptr)) return rc;
            if (random_mode == 1) e->calls += e->n;
        }
        e->hits_valid = false;
        return 0;
    }
    // Graph replay: one captured step = {superpose, project, [random, counter bump], memset, scan, finish}.
    // The captured ED pass reads the post-integrate buffer (crd2), so make that the current one first;
    // mode 0 needs the random-term planes zeroed once, outside the graph.
    if (!e->latest_in_crd2) {
        const size_t off = (size_t)e->t0 * 3 * e->S, cnt = (size_t)(e->t1 - e->t0) * 3 * e->S;
        if (cnt) CK(cudaMemcpyAsync(e->crd2 + off, e->crd + off, sizeof(double) * cnt, cudaMemcpyDeviceToDevice, e->st));
        e->latest_in_crd2 = true;
    }
    if (random_mode == 0) { if (int rc = do_random(e, 0, 0, 0, 0)) return rc; }
    if (!e->d_calls) CK(dev_alloc(&e->d_calls, 1));
    CK(cudaMemcpyAsync(e->d_calls, &e->calls, sizeof(long), cudaMemcpyHostToDevice, e->st));
    if (!e->graph_valid || e->graph_random_mode != random_mode) {
        drop_graph(e);
        // tables that depend only on the parameters / the pair slice: built here, NOT inside the capture
        // (drop_graph has just invalidated the second one)
        if (random_mode == 1) if (int rc = ensure_noise(e)) return rc;
        if (e->scan_mode == 4) if (int rc = ensure_need(e)) return rc;
        CK(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
        const bool zero_flag = e->rnd_zero;
        const long long counted = e->launches;
        int rc = one_step(e, random_mode, e->d_calls);
        e->graph_launches = e->launches - counted;   // kernels in one captured step (nothing ran yet)
        e->launches = counted;
        cudaError_t end = cudaStreamEndCapture(e->st, &e->graph);
        e->rnd_zero = random_mode == 0 ? zero_flag : false;
        if (rc) { if (e->graph) { cudaGraphDestroy(e->graph); e->graph = nullptr; } return rc; }
        CK(end);
        // a rebuilt pair list changes kernel arguments and grid sizes but rarely the shape of the graph:
        // patch the executable graph in place when that works, instantiate a new one otherwise
        bool updated = false;
        if (e->graph_exec) {
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(e->graph_exec, e->graph, &info) == cudaSuccess) updated = true;
            else { cudaGetLastError(); cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
        }
        if (!updated) CK(cudaGraphInstantiate(&e->graph_exec, e->graph, 0));
        e->graph_valid = true; e->graph_random_mode = random_mode;
        e->latest_in_crd2 = true;
    }
    for (int it = 0; it < nsteps; it++) CK(cudaGraphLaunch(e->graph_exec, e->st));
    e->launches += (long long)nsteps * e->graph_launches;
    if (random_mode == 1) { e->calls += (long)nsteps * e->n; e->rnd_zero = false; }
    e->latest_in_crd2 = true;
    e->hits_valid = false;
    return 0;
}

// One call = upload the caller's state, run the steps, read the state back (what Master does around
// calculate_Forces with its host Tetrad array, master.cpp:268-343).  Because the host buffers are
// only touched inside the call, the transfers are pipelined with the kernels: coordinates go up
// first, the velocities follow on a second stream while the ED and distance kernels (which do not
// read them) run, and the finish kernel waits for them.
int edmd_step_host(edmd_engine* e, edmd_tetrad* tets, int n, int nsteps, int random_mode) {
    if (int rc = check_ready(e, true)) return rc;
    if (!tets || n != e->n) return fail(EDMD_ERR_ARG, "edmd_step_host: n=%d but engine holds %d tetrads", n, e->n);
    if (nsteps < 1) return fail(EDMD_ERR_ARG, "edmd_step_host: nsteps must be >= 1");
    if (e->peer_mode) return fail(EDMD_ERR_STATE, "edmd_step_host runs on a single-GPU engine; with peers attached use "
                                  "edmd_set_state / edmd_step / edmd_peer_gather / edmd_get_state");
    if (random_mode != 0 && random_mode != 1) return fail(EDMD_ERR_ARG, "random mode %d unknown", random_mode);
    const bool pipelined = dense_field(e, tets, 0) && dense_field(e, tets, 1) && e->nb_clip && !e->timing;
    if (!pipelined) {
        if (int rc = edmd_set_state(e, tets, n)) return rc;
        if (int rc = edmd_step(e, nsteps, random_mode)) return rc;
        return edmd_get_state(e, tets, n);
    }
    CK(cudaSetDevice(e->device));
    if (!e->st_copy) {
        CK(cudaStreamCreateWithFlags(&e->st_copy, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&e->ev_crd_in, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->ev_vel_in, cudaEventDisableTiming));
    }
    e->latest_in_crd2 = false;                       // the device state is replaced as a whole
    if (int rc = snap_fence(e)) return rc;
    const size_t bytes = sizeof(double) * e->total3;
    CK(cudaMemcpyAsync(e->d_stage, tets[0].coordinates, bytes, cudaMemcpyHostToDevice, e->st));
    CK(cudaEventRecord(e->ev_crd_in, e->st));
    CK(launch_pack_planes(e->d_stage, e->d_off3, e->d_natoms, e->crd, e->n, e->S, e->st));
    CK(cudaStreamWaitEvent(e->st_copy, e->ev_crd_in, 0));          // one upload at a time on the link
    CK(cudaMemcpyAsync(e->d_stage + e->total3, tets[0].velocities, bytes, cudaMemcpyHostToDevice, e->st_copy));
    CK(launch_pack_planes(e->d_stage + e->total3, e->d_off3, e->d_natoms, e->vel, e->n, e->S, e->st_copy));
    CK(cudaEventRecord(e->ev_vel_in, e->st_copy));
    e->launches += 2;
    e->hits_valid = false;
    // first step as separate launches so that the finish kernel alone waits for the velocities
    if (int rc = do_ed(e)) return rc;
    if (int rc = do_random(e, random_mode, e->time0, e->rank, e->calls, nullptr)) return rc;
    if (int rc = do_scan(e)) return rc;
    CK(cudaStreamWaitEvent(e->st, e->ev_vel_in, 0));
    if (int rc = do_finish(e)) return rc;
    if (random_mode == 1) e->calls += e->n;
    e->hits_valid = false;
    if (nsteps > 1) if (int rc = edmd_step(e, nsteps - 1, random_mode)) return rc;
    return edmd_get_state(e, tets, n);
}

// Asynchronous output (north_star: "trajectory output stays host-side on a side stream"; replaces the blocking
// copies around Master::write_Info / write_Crds, master.cpp:388-419).  The state is unpacked to the reference
// layout in stream order (so it is the state at this point of the run), then drains to the caller's buffers on
// a second stream while the engine stream goes on stepping.
int edmd_snapshot_begin(edmd_engine* e, int ticket, double* crd, double* vel, double* obs) {
    if (int rc = check_ready(e, false)) return rc;
    if (ticket < 0 || ticket > 1) return fail(EDMD_ERR_ARG, "snapshot ticket must be 0 or 1");
    if (e->snap_pending[ticket]) return fail(EDMD_ERR_STATE, "snapshot %d is still in flight: call edmd_snapshot_wait first", ticket);
    CK(cudaSetDevice(e->device));
    if (!e->st_out) {
        CK(cudaStreamCreateWithFlags(&e->st_out, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&e->ev_staged, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->ev_snap[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->ev_snap[1], cudaEventDisableTiming));
    }
    if (int rc = normalize(e)) return rc;
    if (int rc = snap_fence(e)) return rc;      // the other ticket may still be reading the staging buffer
    if (crd) { CK(launch_unpack_planes(e->crd, e->d_off3, e->d_natoms, e->d_stage, e->n, e->S, e->st)); e->launches++; }
    if (vel) { CK(launch_unpack_planes(e->vel, e->d_off3, e->d_natoms, e->d_stage + e->total3, e->n, e->S, e->st)); e->launches++; }
    if (obs) {
        if (!e->d_obs_snap) CK(dev_alloc(&e->d_obs_snap, 4 * (size_t)e->n));
        CK(cudaMemcpyAsync(e->d_obs_snap, e->obs, sizeof(double) * 4 * e->n, cudaMemcpyDeviceToDevice, e->st));
    }
    CK(cudaEventRecord(e->ev_staged, e->st));
    CK(cudaStreamWaitEvent(e->st_out, e->ev_staged, 0));
    if (crd) CK(cudaMemcpyAsync(crd, e->d_stage, sizeof(double) * e->total3, cudaMemcpyDeviceToHost, e->st_out));
    if (vel) CK(cudaMemcpyAsync(vel, e->d_stage + e->total3, sizeof(double) * e->total3, cudaMemcpyDeviceToHost, e->st_out));
    if (obs) CK(cudaMemcpyAsync(obs, e->d_obs_snap, sizeof(double) * 4 * e->n, cudaMemcpyDeviceToHost, e->st_out));
    CK(cudaEventRecord(e->ev_snap[ticket], e->st_out));
    e->snap_pending[ticket] = true;
    return 0;
}

int edmd_snapshot_wait(edmd_engine* e, int ticket) {
    if (!e || ticket < 0 || ticket > 1) return fail(EDMD_ERR_ARG, "bad snapshot ticket");
    if (!e->snap_pending[ticket]) return 0;
    CK(cudaEventSynchronize(e->ev_snap[ticket]));
    e->snap_pending[ticket] = false;
    return 0;
}

long long edmd_state_doubles(edmd_engine* e) { return e ? e->total3 : 0; }

int edmd_sync(edmd_engine* e) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->st));
    return peer_check(e);
}

int edmd_device_buffer(edmd_engine* e, int which, void** dptr, size_t* bytes) {
    if (int rc = check_ready(e, false)) return rc;
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    const size_t planes = sizeof(double) * (size_t)e->n * 3 * e->S;
    void* p = nullptr; size_t b = 0;
    switch (which) {
        case EDMD_BUF_CRD: p = e->crd; b = planes; break;
        case EDMD_BUF_VEL: p = e->vel; b = planes; break;
        case EDMD_BUF_FED: p = e->fed; b = planes; break;
        case EDMD_BUF_RND: p = e->rnd; b = planes; break;
        case EDMD_BUF_FNB: p = e->fnb; b = planes; break;
        case EDMD_BUF_HIT: p = cur_hits(e); b = sizeof(unsigned long long) * (size_t)e->np; break;
        case EDMD_BUF_CENTROID: p = e->cen; b = sizeof(double) * 4 * (size_t)e->n; break;
        default: return fail(EDMD_ERR_ARG, "unknown buffer %d", which);
    }
    if (dptr) *dptr = p;
    if (bytes) *bytes = b;
    return 0;
}

int edmd_atom_stride(edmd_engine* e) { return e ? e->S : 0; }
int edmd_num_tetrads(edmd_engine* e) { return e ? e->n : 0; }
long long edmd_num_pairs(edmd_engine* e) { return e ? e->np : 0; }
void* edmd_stream(edmd_engine* e) { return e ? (void*)e->st : nullptr; }

int edmd_timer_start(edmd_engine* e) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    if (!e->span0) { CK(cudaEventCreate(&e->span0)); CK(cudaEventCreate(&e->span1)); }
    CK(cudaEventRecord(e->span0, e->st));
    return 0;
}
int edmd_timer_stop(edmd_engine* e, double* ms) {
    if (!e || !e->span0) return fail(EDMD_ERR_STATE, "edmd_timer_stop without edmd_timer_start");
    CK(cudaSetDevice(e->device));
    CK(cudaEventRecord(e->span1, e->st));
    CK(cudaEventSynchronize(e->span1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, e->span0, e->span1));
    if (ms) *ms = f;
    return 0;
}

int edmd_mark(edmd_engine* e, int slot) {
    if (!e || slot < 0 || slot > 65535) return fail(EDMD_ERR_ARG, "bad mark slot");
    CK(cudaSetDevice(e->device));
    while ((int)e->marks.size() <= slot) { cudaEvent_t ev; CK(cudaEventCreate(&ev)); e->marks.push_back(ev); }
    CK(cudaEventRecord(e->marks[slot], e->st));
    return 0;
}
int edmd_mark_elapsed(edmd_engine* e, int slot_a, int slot_b, double* ms) {
    if (!e || slot_a < 0 || slot_b < 0 || slot_a >= (int)e->marks.size() || slot_b >= (int)e->marks.size())
        return fail(EDMD_ERR_ARG, "mark slot not recorded");
    CK(cudaSetDevice(e->device));
    CK(cudaEventSynchronize(e->marks[slot_b]));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, e->marks[slot_a], e->marks[slot_b]));
    if (ms) *ms = f;
    return 0;
}

int edmd_timing_enable(edmd_engine* e, int on) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    e->timing = on != 0;
    return 0;
}
int edmd_timing_get(edmd_engine* e, int which, double* total_ms, long long* launches) {
    if (!e || which < 0 || which >= T_COUNT) return fail(EDMD_ERR_ARG, "bad timer id");
    CK(cudaSetDevice(e->device));
    timer_fold(e, which);
    if (total_ms) *total_ms = e->timers[which].total_ms;
    if (launches) *launches = e->timers[which].launches;
    return 0;
}
int edmd_timing_reset(edmd_engine* e) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    for (int i = 0; i < T_COUNT; i++) { timer_fold(e, i); e->timers[i].total_ms = 0.0; e->timers[i].launches = 0; }
    return 0;
}
long long edmd_launch_count(edmd_engine* e) { return e ? e->launches : 0; }

int edmd_nb_stats(edmd_engine* e, long long* flagged_pairs) {
    if (int rc = check_ready(e, true)) return rc;
    CK(cudaSetDevice(e->device));
    std::vector<unsigned long long> h((size_t)e->np);
    CK(cudaStreamSynchronize(e->st));
    if (e->np) CK(cudaMemcpy(h.data(), cur_hits(e), sizeof(unsigned long long) * (size_t)e->np, cudaMemcpyDeviceToHost));
    long long c = 0;
    for (long long p = e->p0; p < e->p1; p++) c += h[p] != 0;
    if (flagged_pairs) *flagged_pairs = c;
    return 0;
}

int edmd_nb_stats_ex(edmd_engine* e, long long out[5]) {
    if (int rc = check_ready(e, true)) return rc;
    if (!out) return fail(EDMD_ERR_ARG, "null output");
    CK(cudaSetDevice(e->device));
    std::vector<unsigned long long> h((size_t)e->np);
    std::vector<int2> pr((size_t)e->np);
    CK(cudaStreamSynchronize(e->st));
    if (e->np) {
        CK(cudaMemcpy(h.data(), cur_hits(e), sizeof(unsigned long long) * (size_t)e->np, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pr.data(), e->d_pairs, sizeof(int2) * (size_t)e->np, cudaMemcpyDeviceToHost));
    }
    std::vector<unsigned char> marked((size_t)e->n, 0);
    out[0] = out[1] = out[2] = out[3] = out[4] = 0;
    for (long long p = 0; p < e->np; p++) {
        if (!h[p]) continue;
        out[0]++;
        out[2] += __builtin_popcountll(h[p]);
        marked[pr[p].x] = 1; marked[pr[p].y] = 1;
    }
    for (int t = e->t0; t < e->t1; t++) out[1] += marked[t];
    if (e->scan_mode == 4 && e->d_pmask) {
        unsigned cnt[2] = {0, 0};   // sparse-mask items from the front of the work buffer, dense-mask items from its back
        CK(cudaMemcpy(cnt, e->d_mark + e->n + 2, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost));
        out[3] = (long long)cnt[0] + cnt[1];
        // work items {pair index, sphere mask} (nb.cu: CullItem)
        std::vector<unsigned long long> items(2 * (size_t)std::max(cnt[0], cnt[1]));
        if (cnt[0]) CK(cudaMemcpy(items.data(), e->d_pmask, sizeof(unsigned long long) * 2 * cnt[0], cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt[0]; i++) out[4] += __builtin_popcountll(items[2 * i + 1]);
        if (cnt[1]) CK(cudaMemcpy(items.data(), e->d_pmask + 2 * ((size_t)e->pmask_cap - cnt[1]), sizeof(unsigned long long) * 2 * cnt[1], cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt[1]; i++) out[4] += __builtin_popcountll(items[2 * i + 1]);
    }
    return 0;
}

void* edmd_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { fail(EDMD_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes); return nullptr; }
    return p;
}
void edmd_free_pinned(void* p) { if (p) cudaFreeHost(p); }

// ---- multi-GPU: attaching peers ---------------------------------------------------------------------
namespace {

struct PeerPtrs { double* crd; double* vel; unsigned long long* hitx; unsigned long long* flags; double* obs; };

// allocate (or keep) the peer-visible buffers of this engine and reset the barrier state
int peer_prepare(edmd_engine* e) {
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->st));
    close_peers(e);
    const long long need = std::max<long long>(std::max<long long>(2 * e->np, 16LL * e->n), 4096);
    if (need > e->hitx_cap) {
        cudaFree(e->d_hitx); e->d_hitx = nullptr; e->hitx_cap = 0;
        CK(cudaMalloc((void**)&e->d_hitx, sizeof(unsigned long long) * (size_t)need));
        e->hitx_cap = need;
    }
    if (!e->d_flags) {
        CK(cudaMalloc((void**)&e->d_flags, sizeof(unsigned long long) * 32));
        CK(cudaMalloc((void**)&e->d_epoch, sizeof(unsigned long long)));
        CK(cudaMalloc((void**)&e->d_peer_err, sizeof(int)));
    }
    CK(cudaMemset(e->d_hitx, 0, sizeof(unsigned long long) * (size_t)e->hitx_cap));
    CK(cudaMemset(e->d_flags, 0, sizeof(unsigned long long) * 32));
    CK(cudaMemset(e->d_epoch, 0, sizeof(unsigned long long)));
    CK(cudaMemset(e->d_peer_err, 0, sizeof(int)));
    if (const char* t = getenv("EDMD_PEER_TIMEOUT_S")) {
        const double sec = atof(t);
        if (sec > 0.0) e->peer_timeout_ns = (unsigned long long)(sec * 1e9);
    }
    return 0;
}

// install the tables of mapped pointers (own rank: the local buffers)
int peer_install(edmd_engine* e, const std::vector<PeerPtrs>& all, int world, int rank) {
    std::vector<double*> crd(world);
    std::vector<unsigned long long*> hit(world), flg(world);
    e->peer_crd.assign(world, nullptr); e->peer_vel.assign(world, nullptr); e->peer_obs.assign(world, nullptr);
    for (int r = 0; r < world; r++) {
        crd[r] = all[r].crd; hit[r] = all[r].hitx; flg[r] = all[r].flags;
        e->peer_crd[r] = all[r].crd; e->peer_vel[r] = all[r].vel; e->peer_obs[r] = all[r].obs;
    }
    CK(cudaMalloc((void**)&e->d_crd_tbl, sizeof(double*) * world));
    CK(cudaMalloc((void**)&e->d_hit_tbl, sizeof(unsigned long long*) * world));
    CK(cudaMalloc((void**)&e->d_flag_tbl, sizeof(unsigned long long*) * world));
    CK(cudaMemcpy(e->d_crd_tbl, crd.data(), sizeof(double*) * world, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_hit_tbl, hit.data(), sizeof(unsigned long long*) * world, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_flag_tbl, flg.data(), sizeof(unsigned long long*) * world, cudaMemcpyHostToDevice));
    e->world = world; e->rank_id = rank; e->peer_mode = true;
    e->peer_replicated = (e->t0 == 0 && e->t1 == e->n);
    drop_graph(e);
    return rebuild_halo(e);
}

// every rank's owned coordinates / velocities / observables -> this rank's buffers (contiguous owner blocks:
// plain device-to-device copies through the mappings).  Collective: barrier before (owners have finished
// writing) and after (nobody overwrites what a peer is still copying).
int peer_gather(edmd_engine* e, int what) {
    if (!e->peer_mode) return 0;
    if (int rc = normalize(e)) return rc;
    if (int rc = peer_barrier(e)) return rc;
    if (!e->peer_replicated) {
        const size_t w3 = 3 * (size_t)e->S;
        for (int r = 0; r < e->world; r++) {
            if (r == e->rank_id) continue;
            int a, b; owned_range(e, r, &a, &b);
            if (b <= a) continue;
            const size_t off = (size_t)a * w3, cnt = (size_t)(b - a) * w3;
            if (what & 1) CK(cudaMemcpyAsync(e->crd + off, e->peer_crd[r] + off, sizeof(double) * cnt, cudaMemcpyDefault, e->st));
            if (what & 2) CK(cudaMemcpyAsync(e->vel + off, e->peer_vel[r] + off, sizeof(double) * cnt, cudaMemcpyDefault, e->st));
            if (what & 4) {
                const size_t n = (size_t)e->n;
                CK(cudaMemcpyAsync(e->obs + a, e->peer_obs[r] + a, sizeof(double) * (b - a), cudaMemcpyDefault, e->st));
                CK(cudaMemcpyAsync(e->obs + n + 2 * (size_t)a, e->peer_obs[r] + n + 2 * (size_t)a, sizeof(double) * 2 * (b - a), cudaMemcpyDefault, e->st));
                CK(cudaMemcpyAsync(e->obs + 3 * n + a, e->peer_obs[r] + 3 * n + a, sizeof(double) * (b - a), cudaMemcpyDefault, e->st));
            }
        }
    }
    return peer_barrier(e);
}

}  // namespace

/* blob = five CUDA IPC handles: coordinates, velocities, mask buffer, barrier flags, observables */
int edmd_peer_export(edmd_engine* e, unsigned char* blob) {
    if (int rc = check_ready(e, false)) return rc;
    if (!blob) return fail(EDMD_ERR_ARG, "null blob");
    if (int rc = normalize(e)) return rc;
    if (int rc = peer_prepare(e)) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    static_assert(EDMD_PEER_BLOB_BYTES == 5 * 64, "blob layout");
    void* bufs[5] = {e->crd, e->vel, e->d_hitx, e->d_flags, e->obs};
    for (int k = 0; k < 5; k++) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, bufs[k]));
        memcpy(blob + 64 * k, &h, 64);
    }
    return 0;
}

int edmd_peer_attach(edmd_engine* e, const unsigned char* blobs, int world, int rank) {
    if (int rc = check_ready(e, false)) return rc;
    if (!blobs || world < 1 || world > 32 || rank < 0 || rank >= world || !e->d_hitx)
        return fail(EDMD_ERR_ARG, "edmd_peer_attach: bad arguments (call edmd_peer_export first; world <= 32)");
    CK(cudaSetDevice(e->device));
    std::vector<PeerPtrs> all(world);
    for (int r = 0; r < world; r++) {
        if (r == rank) { all[r] = PeerPtrs{e->crd, e->vel, e->d_hitx, e->d_flags, e->obs}; continue; }
        void* p[5];
        for (int k = 0; k < 5; k++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, blobs + (size_t)r * EDMD_PEER_BLOB_BYTES + 64 * k, 64);
            p[k] = nullptr;
            CK(cudaIpcOpenMemHandle(&p[k], h, cudaIpcMemLazyEnablePeerAccess));
            e->peer_opened.push_back(p[k]);
        }
        all[r] = PeerPtrs{(double*)p[0], (double*)p[1], (unsigned long long*)p[2], (unsigned long long*)p[3], (double*)p[4]};
    }
    return peer_install(e, all, world, rank);
}

int edmd_peer_detach(edmd_engine* e) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->st));
    const int rc = peer_check(e);
    close_peers(e);
    drop_graph(e);
    return rc;
}

int edmd_peer_gather(edmd_engine* e, int what) {
    if (int rc = check_ready(e, false)) return rc;
    if (what < 1 || what > 7) return fail(EDMD_ERR_ARG, "edmd_peer_gather: what must be a combination of 1 (coordinates), 2 (velocities), 4 (observables)");
    CK(cudaSetDevice(e->device));
    return peer_gather(e, what);
}

int edmd_world(edmd_engine* e, int* world, int* rank) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    if (world) *world = e->peer_mode ? e->world : 1;
    if (rank) *rank = e->peer_mode ? e->rank_id : 0;
    return 0;
}

int edmd_measure_fp64_peak(int device, double* tflops, double* sm_mhz_est) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(EDMD_ERR_CUDA, "no CUDA device available");
    CK(cudaSetDevice(device));
    CK(measure_fp64_peak(tflops, sm_mhz_est));
    return 0;
}

// ---- single process, several GPUs ------------------------------------------------------------------
}  // extern "C" (reopened below)

struct edmd_group {
    std::vector<edmd_engine*> eng;
    bool attached = false;
    int replicate_max = 0;     // systems up to this size run the per-tetrad kernels on every GPU (0: never)
};

namespace {

int group_attach(edmd_group* g) {
    const int world = (int)g->eng.size();
    for (edmd_engine* e : g->eng) {
        CK(cudaSetDevice(e->device));
        if (int rc = normalize(e)) return rc;
        if (int rc = peer_prepare(e)) return rc;
    }
    std::vector<PeerPtrs> all(world);
    for (int r = 0; r < world; r++) {
        edmd_engine* e = g->eng[r];
        all[r] = PeerPtrs{e->crd, e->vel, e->d_hitx, e->d_flags, e->obs};
    }
    for (int r = 0; r < world; r++) {
        CK(cudaSetDevice(g->eng[r]->device));
        if (int rc = peer_install(g->eng[r], all, world, r)) return rc;
    }
    g->attached = true;
    return 0;
}

int group_detach(edmd_group* g) {
    for (edmd_engine* e : g->eng) {
        CK(cudaSetDevice(e->device));
        CK(cudaStreamSynchronize(e->st));
    }
    for (edmd_engine* e : g->eng) { close_peers(e); drop_graph(e); }
    g->attached = false;
    return 0;
}

// enqueue the collective on every GPU before anything waits for one of them
int group_gather(edmd_group* g, int what) {
    if (!g->attached) return 0;
    for (edmd_engine* e : g->eng) {
        CK(cudaSetDevice(e->device));
        if (int rc = peer_gather(e, what)) return rc;
    }
    return 0;
}

}  // namespace

extern "C" {

int edmd_group_create(const edmd_params* p, int n_gpus, const int* devices, edmd_group** out) {
    if (!out) return fail(EDMD_ERR_ARG, "out is null");
    *out = nullptr;
    if (n_gpus < 1 || n_gpus > 32) return fail(EDMD_ERR_ARG, "n_gpus = %d out of range [1,32]", n_gpus);
    edmd_group* g = new edmd_group();
    for (int r = 0; r < n_gpus; r++) {
        edmd_engine* e = nullptr;
        if (int rc = edmd_create(p, devices ? devices[r] : r, &e)) { edmd_group_destroy(g); return rc; }
        g->eng.push_back(e);
    }
    // every GPU reads and writes every other GPU's buffers directly (NVLink / NVSwitch)
    for (int a = 0; a < n_gpus; a++)
        for (int b = 0; b < n_gpus; b++) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->eng[a]->device, g->eng[b]->device);
            if (!can) {
                const int da = g->eng[a]->device, db = g->eng[b]->device;
                edmd_group_destroy(g);
                return fail(EDMD_ERR_CUDA, "device %d cannot access device %d's memory: no peer path", da, db);
            }
            cudaSetDevice(g->eng[a]->device);
            cudaError_t err = cudaDeviceEnablePeerAccess(g->eng[b]->device, 0);
            if (err == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); err = cudaSuccess; }
            if (err != cudaSuccess) {
                edmd_group_destroy(g);
                return fail(EDMD_ERR_CUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(err));
            }
        }
    *out = g;
    return 0;
}

void edmd_group_destroy(edmd_group* g) {
    if (!g) return;
    if (g->attached) group_detach(g);
    for (edmd_engine* e : g->eng) edmd_destroy(e);
    delete g;
}

int edmd_group_size(edmd_group* g) { return g ? (int)g->eng.size() : 0; }
edmd_engine* edmd_group_engine(edmd_group* g, int rank) {
    return (g && rank >= 0 && rank < (int)g->eng.size()) ? g->eng[rank] : nullptr;
}

int edmd_group_upload_tetrads(edmd_group* g, const edmd_tetrad* tets, int n, const int* bp_atoms) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    if (g->attached) if (int rc = group_detach(g)) return rc;
    for (edmd_engine* e : g->eng) if (int rc = edmd_upload_tetrads(e, tets, n, bp_atoms)) return rc;
    return 0;
}

int edmd_group_set_state(edmd_group* g, const edmd_tetrad* tets, int n) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    for (edmd_engine* e : g->eng) if (int rc = edmd_set_state(e, tets, n)) return rc;
    return 0;
}

int edmd_group_get_state(edmd_group* g, edmd_tetrad* tets, int n) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    if (int rc = group_gather(g, 3)) return rc;
    return edmd_get_state(g->eng[0], tets, n);
}

int edmd_group_set_rng(edmd_group* g, long time0, int rank, long calls_before) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    for (edmd_engine* e : g->eng) if (int rc = edmd_set_rng(e, time0, rank, calls_before)) return rc;
    return 0;
}

int edmd_group_build_pairlist(edmd_group* g, long long* num_pairs) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    const int world = (int)g->eng.size();
    if (int rc = group_gather(g, 1)) return rc;        // the list is built from everybody's coordinates
    long long np = -1;
    bool regrow = false;
    for (edmd_engine* e : g->eng) {
        long long mine = 0;
        const int rc = edmd_build_pairlist(e, &mine);
        if (rc == EDMD_ERR_STATE && g->attached) { regrow = true; break; }
        if (rc) return rc;
        if (np >= 0 && mine != np) return fail(EDMD_ERR_STATE, "GPUs disagree on the pair list: %lld vs %lld pairs", mine, np);
        np = mine;
    }
    if (regrow) {   // the list outgrew the peer-visible mask buffers: map larger ones
        if (int rc = group_detach(g)) return rc;
        np = -1;
        for (edmd_engine* e : g->eng) {
            long long mine = 0;
            if (int rc = edmd_build_pairlist(e, &mine)) return rc;
            np = mine;
        }
    }
    if (world > 1 && !g->attached) if (int rc = group_attach(g)) return rc;
    // work split: contiguous tetrad blocks, equal pair slices (generate_Indexes, master.cpp:214-242)
    const int n = g->eng[0]->n;
    const bool replicate = n <= g->replicate_max;
    const int tchunk = (n + world - 1) / world;
    const long long pchunk = (np + world - 1) / world;
    for (int r = 0; r < world; r++) {
        const int t0 = replicate ? 0 : std::min(n, r * tchunk), t1 = replicate ? n : std::min(n, t0 + tchunk);
        const long long p0 = std::min(np, r * pchunk), p1 = std::min(np, p0 + pchunk);
        if (int rc = edmd_set_shard(g->eng[r], t0, t1, p0, p1)) return rc;
    }
    if (num_pairs) *num_pairs = np;
    return 0;
}

int edmd_group_step(edmd_group* g, int nsteps, int random_mode) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    for (edmd_engine* e : g->eng) if (int rc = edmd_step(e, nsteps, random_mode)) return rc;
    return 0;
}

int edmd_group_merge(edmd_group* g) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    if (int rc = group_gather(g, 3)) return rc;
    for (edmd_engine* e : g->eng) if (int rc = edmd_merge(e)) return rc;
    return 0;
}

int edmd_group_get_energies(edmd_group* g, double out[4]) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    if (int rc = group_gather(g, 4)) return rc;
    return edmd_get_energies(g->eng[0], out);
}

int edmd_group_snapshot_begin(edmd_group* g, int ticket, double* crd, double* vel, double* obs) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    if (int rc = group_gather(g, (crd ? 1 : 0) | (vel ? 2 : 0) | (obs ? 4 : 0))) return rc;
    return edmd_snapshot_begin(g->eng[0], ticket, crd, vel, obs);
}
int edmd_group_snapshot_wait(edmd_group* g, int ticket) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    return edmd_snapshot_wait(g->eng[0], ticket);
}

int edmd_group_sync(edmd_group* g) {
    if (!g) return fail(EDMD_ERR_ARG, "null group");
    for (edmd_engine* e : g->eng) if (int rc = edmd_sync(e)) return rc;
    return 0;
}

}  // extern "C"
