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
#include <cstdarg>
#include <cstdio>
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

enum TimerId { T_ED = 0, T_SCAN, T_ACC, T_RNG, T_MERGE, T_PAIRS, T_INT, T_LAYOUT, T_COUNT };

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
    double *e_ed = nullptr, *e_nb = nullptr, *temp = nullptr, *cen = nullptr, *sup = nullptr;
    int* d_mark = nullptr;            // [2n+1] per-tetrad "in a flagged pair" flags, list counter, list of marked tetrads
    unsigned char* d_dirty = nullptr; // [n] fnb of this tetrad is non-zero from an earlier step
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

    // peer-shared hit flags (fused scan + broadcast over NVLink): two buffers used by step parity,
    // exported with CUDA IPC; peer_tbl[b][r] = rank r's buffer b as mapped into this process
    unsigned long long* d_hitx[2] = {nullptr, nullptr};
    long long hitx_cap = 0;
    unsigned long long** d_peer_tbl = nullptr;   // device array [2][world]
    std::vector<void*> peer_opened;
    int world = 1, rank_id = 0, hit_parity = 0;
    bool peer_mode = false;

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

    int ed_kernel = 0;   // projection kernel: 0 automatic, 1 thread-per-atom, 2 warp-per-tetrad (edmd_set_ed_kernel)
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
    double* d_pspre = nullptr;              // [P][4] per-set superposition constants (ed.cu)
    double* d_noise = nullptr; bool noise_valid = false;   // [P][3][S] Langevin noise amplitudes
    cudaStream_t st_copy = nullptr;                 // edmd_step_host: velocity upload beside the ED / scan kernels
    cudaEvent_t ev_crd_in = nullptr, ev_vel_in = nullptr;
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
    close_peers(e);
    cudaFree(e->d_hitx[0]); cudaFree(e->d_hitx[1]); e->d_hitx[0] = e->d_hitx[1] = nullptr; e->hitx_cap = 0;
    cudaFree(e->d_calls); e->d_calls = nullptr;
    cudaFree(e->d_mark); cudaFree(e->d_dirty); e->d_mark = nullptr; e->d_dirty = nullptr;
    cudaFree(e->d_perm); e->d_perm = nullptr;
    cudaFree(e->d_grp); e->d_grp = nullptr;
    cudaFree(e->d_gb); cudaFree(e->d_pmask); e->d_gb = nullptr; e->d_pmask = nullptr; e->pmask_cap = 0;
    cudaFree(e->d_noise); e->d_noise = nullptr; e->noise_valid = false;
    cudaFree(e->d_pspre); e->d_pspre = nullptr;
    cudaFree(e->d_need); e->d_need = nullptr; e->need_valid = false;
    cudaFree(e->d_order); e->d_order = nullptr;
    cudaFree(e->d_cell); cudaFree(e->d_cell_sorted); cudaFree(e->d_tet); cudaFree(e->d_tet_sorted);
    cudaFree(e->d_partners); cudaFree(e->d_partners_sorted); cudaFree(e->d_cub_tmp);
    e->d_cell = e->d_cell_sorted = nullptr; e->d_tet = e->d_tet_sorted = e->d_partners = e->d_partners_sorted = nullptr;
    e->d_cub_tmp = nullptr; e->cub_tmp_bytes = 0; e->partners_cap = 0;
    cudaFree(e->d_natoms); cudaFree(e->d_param_id); cudaFree(e->d_ps_na); cudaFree(e->d_ps_nev);
    cudaFree(e->d_off3); cudaFree(e->d_avg); cudaFree(e->d_mass); cudaFree(e->d_chg);
    cudaFree(e->d_eval); cudaFree(e->d_evec);
    cudaFree(e->crd); cudaFree(e->crd2); cudaFree(e->vel); cudaFree(e->fed); cudaFree(e->rnd); cudaFree(e->fnb);
    cudaFree(e->e_ed); cudaFree(e->e_nb); cudaFree(e->temp); cudaFree(e->cen); cudaFree(e->sup);
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
    close_peers(e);   // the driver re-establishes the peer-shared hit buffers after every rebuild
    if (np <= e->pairs_cap) return 0;
    long long cap = std::max(np, e->pairs_cap * 3 / 2 + 1024);
    cudaFree(e->d_pairs); cudaFree(e->d_hit); cudaFree(e->d_pmask);
    e->d_pairs = nullptr; e->d_hit = nullptr; e->d_pmask = nullptr; e->pairs_cap = 0; e->pmask_cap = 0;
    CK(dev_alloc(&e->d_pairs, (size_t)cap));
    CK(dev_alloc(&e->d_hit, (size_t)cap));
    CK(dev_alloc(&e->d_pmask, 2 * (size_t)cap + 2));   // counter slot + one {pair, mask} item per pair (nb.cu)
    e->pmask_cap = cap;
    e->pairs_cap = cap;
    return 0;
}

// CSR over tetrads of the pair list, entries in pair-list order (worker.cpp:166-179 accumulates
// in that order).  Host-side: runs every ntsync steps only.
int build_adjacency(edmd_engine* e, const std::vector<int2>& pairs) {
    const int n = e->n;
    const long long np = (long long)pairs.size();
    std::vector<long long> off(n + 1, 0);
    for (long long p = 0; p < np; p++) { off[pairs[p].x + 1]++; off[pairs[p].y + 1]++; }
    for (int t = 0; t < n; t++) off[t + 1] += off[t];
    std::vector<int> partner((size_t)(2 * np) + 1), pidx((size_t)(2 * np) + 1);
    std::vector<long long> cur(off.begin(), off.end() - 1);
    for (long long p = 0; p < np; p++) {
        const int a = pairs[p].x, b = pairs[p].y;
        partner[cur[a]] = b;                       pidx[cur[a]++] = (int)p;
        partner[cur[b]] = a | (int)0x80000000;     pidx[cur[b]++] = (int)p;
    }
    if (2 * np > e->adj_cap) {
        cudaFree(e->d_adj_partner); cudaFree(e->d_adj_pair);
        e->d_adj_partner = e->d_adj_pair = nullptr; e->adj_cap = 0;
        const long long cap = 2 * np + 1024;
        CK(dev_alloc(&e->d_adj_partner, (size_t)cap));
        CK(dev_alloc(&e->d_adj_pair, (size_t)cap));
        e->adj_cap = cap;
    }
    CK(cudaMemcpyAsync(e->d_adj_off, off.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, e->st));
    if (np) {
        CK(cudaMemcpyAsync(e->d_adj_partner, partner.data(), sizeof(int) * 2 * np, cudaMemcpyHostToDevice, e->st));
        CK(cudaMemcpyAsync(e->d_adj_pair, pidx.data(), sizeof(int) * 2 * np, cudaMemcpyHostToDevice, e->st));
    }
    CK(cudaStreamSynchronize(e->st));
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
    cudaFree(e->d_peer_tbl); e->d_peer_tbl = nullptr;
    e->peer_mode = false;
}

unsigned long long* cur_hits(edmd_engine* e) { return e->peer_mode ? e->d_hitx[e->hit_parity] : e->d_hit; }

void drop_graph(edmd_engine* e) {
    e->need_valid = false;   // every caller changes the pair list, the shard or the scan mode
    if (e->graph_exec) { cudaGraphExecDestroy(e->graph_exec); e->graph_exec = nullptr; }
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
    const bool warp = e->ed_kernel == 2 || (e->ed_kernel == 0 && (long long)e->n >= 8LL * e->P);
    CK(launch_ed_forces(e->ps, e->d_param_id, src, e->crd, e->fed, e->e_ed, e->sup, e->d_order, e->sms, e->t0,
                        e->t1 - e->t0, e->sp.scaled, warp, e->st));
    timer_end(e, T_ED, 2);
    return 0;
}
// which tetrads the culled distance pass needs bounding spheres for: the ends of this engine's pairs
int ensure_need(edmd_engine* e) {
    if (e->need_valid) return 0;
    if (!e->d_need) CK(dev_alloc(&e->d_need, (size_t)e->n));
    CK(launch_mark_needed(e->d_pairs, e->p0, e->p1, e->d_need, e->n, e->st));
    e->launches++;
    e->need_valid = true;
    return 0;
}
int do_scan(edmd_engine* e) {
    const double cut_eff = nb_effective_cut2(e->sp);
    if (e->scan_mode == 4 && cut_eff >= 1e-3 && cut_eff < 1e6) {
        if (int rc = ensure_need(e)) return rc;
        timer_begin(e, T_SCAN);
        CK(launch_nb_scan_cull(e->crd, e->d_natoms, e->d_pairs, cur_hits(e), e->p0, e->p1, e->n, e->S, cut_eff, e->sms,
                               e->d_gb, e->d_pmask, e->d_param_id, e->d_perm, e->d_need,
                               e->peer_mode ? e->d_peer_tbl + (size_t)e->hit_parity * e->world : nullptr, e->world, e->st));
        timer_end(e, T_SCAN, e->p1 > e->p0 ? 3 : 0);
        e->hits_valid = true;
        return 0;
    }
    timer_begin(e, T_SCAN);
    CK(launch_nb_scan(e->crd, e->d_natoms, e->d_pairs, cur_hits(e), e->p0, e->p1, e->S,
                      nb_effective_cut2(e->sp), e->scan_mode, e->sms,
                      e->peer_mode ? e->d_peer_tbl + (size_t)e->hit_parity * e->world : nullptr, e->world, e->st));
    timer_end(e, T_SCAN, e->p1 > e->p0 ? 1 : 0);
    e->hits_valid = true;
    return 0;
}
int do_accumulate(edmd_engine* e) {
    CK(cudaMemsetAsync(e->d_dirty, 1, (size_t)e->n, e->st));
    timer_begin(e, T_ACC);
    CK(launch_nb_accumulate(e->crd, e->d_natoms, e->d_param_id, e->d_chg, cur_hits(e), e->d_adj_off,
                            e->d_adj_partner, e->d_adj_pair, e->d_perm, e->d_grp, e->fnb, e->e_nb, e->t0,
                            e->t1 - e->t0, e->S, e->sp, e->nb_clip, e->st));
    timer_end(e, T_ACC, 1);
    return 0;
}
// accumulate + clip + velocity + coordinate update in one kernel (edmd_step, edmd_nb_finish)
int do_integrate(edmd_engine* e, int do_vel, int do_crd);
int do_finish(edmd_engine* e) {
    if (int rc = normalize(e)) return rc;
    if (!e->nb_clip) {   // unclipped forces requested: run the two unfused kernels
        if (int rc = do_accumulate(e)) return rc;
        return do_integrate(e, 1, 1);
    }
    timer_begin(e, T_ACC);
    CK(launch_nb_finish_split(e->crd, e->d_natoms, e->d_param_id, e->d_chg, cur_hits(e), e->d_pairs, e->np, e->d_adj_off,
                              e->d_adj_partner, e->d_adj_pair, e->d_perm, e->d_grp, e->fnb, e->e_nb, e->ps, e->vel,
                              e->crd2, e->fed, e->rnd_zero ? nullptr : e->rnd, e->temp, e->d_mark, e->d_dirty, e->n,
                              e->t0, e->t1 - e->t0, e->S, e->sms, e->sp, e->st));
    timer_end(e, T_ACC, 3);
    if (e->peer_mode) {   // this buffer is written again by the peers' scans two steps from now
        CK(cudaMemsetAsync(e->d_hitx[e->hit_parity], 0, sizeof(unsigned long long) * (size_t)e->np, e->st));
        e->hit_parity ^= 1;
    }
    e->latest_in_crd2 = true;
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
int do_random(edmd_engine* e, int mode, long time0, int rank, long calls_before, const long* calls_dev = nullptr) {
    if (mode == 0) {
        if (!e->rnd_zero) {
            CK(cudaMemsetAsync(e->rnd, 0, sizeof(double) * (size_t)e->n * 3 * e->S, e->st));
            e->rnd_zero = true;
        }
        return 0;
    }
    if (mode != 1) return fail(EDMD_ERR_ARG, "random mode %d unknown (0 = off, 1 = reference stream)", mode);
    if (int rc = ensure_noise(e)) return rc;
    timer_begin(e, T_RNG);
    CK(launch_random_terms(e->ps, e->d_noise, e->d_param_id, e->rnd, e->t0, e->t1 - e->t0, e->sp, time0, rank,
                           calls_before + e->t0, calls_dev, e->st));
    timer_end(e, T_RNG, 1);
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

// planes -> interleaved flat on the device (slot of d_stage), then to the pinned slot
int stage_out(edmd_engine* e, const double* planes, int slot) {
    double* d = e->d_stage + (size_t)slot * e->total3;
    CK(launch_unpack_planes(planes, e->d_off3, e->d_natoms, d, e->n, e->S, e->st));
    e->launches++;
    CK(cudaMemcpyAsync(e->h_stage + (size_t)slot * e->total3, d, sizeof(double) * e->total3,
                       cudaMemcpyDeviceToHost, e->st));
    return 0;
}
int stage_in(edmd_engine* e, double* planes, int slot) {
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
    CK(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
    // EDMD::EDMD defaults (edmd.cpp:23-43)
    e->sp = SimParams{0.002 * 20.455, 2.0 / 20.455, 0.2 * 20.455, 300.0, 0.002 * 300.0, 30.0, 10.0, 5.0, 100.0, 0.0};
    if (p) {
        e->sp.dt = p->dt; e->sp.gamma = p->gamma; e->sp.tautp = p->tautp; e->sp.temperature = p->temperature;
        e->sp.scaled = p->scaled; e->sp.mole_cutoff = p->mole_Cutoff; e->sp.atom_cutoff = p->atom_Cutoff;
        e->sp.mole_least = p->mole_Least;
    }
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
    if (which < 0 || which > 2) return fail(EDMD_ERR_ARG, "ED kernel %d unknown (0 automatic, 1 thread per atom, 2 warp per tetrad)", which);
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
    CK(dev_alloc(&e->e_ed, n)); CK(dev_alloc(&e->e_nb, 2 * (size_t)n)); CK(dev_alloc(&e->temp, n));
    CK(dev_alloc(&e->cen, 4 * (size_t)n));
    CK(dev_alloc(&e->d_mark, 2 * (size_t)n + 2)); CK(dev_alloc(&e->d_dirty, (size_t)n));
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
    for (double* buf : {e->vel, e->fed, e->rnd, e->fnb})
        CK(cudaMemsetAsync(buf, 0, sizeof(double) * planes, e->st));
    CK(cudaMemsetAsync(e->e_ed, 0, sizeof(double) * n, e->st));
    CK(cudaMemsetAsync(e->e_nb, 0, sizeof(double) * 2 * n, e->st));
    CK(cudaMemsetAsync(e->temp, 0, sizeof(double) * n, e->st));
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

    // builder choice: the cell list pays off when the cutoff is small against the extent
    bool cells = false;
    CellGridHost grid{};
    if (e->pair_builder != 1 && e->sp.mole_cutoff > 0.0 && e->sp.mole_cutoff < 1e8) {
        std::vector<double> cen(4 * (size_t)n);
        CK(cudaMemcpyAsync(cen.data(), e->cen, sizeof(double) * cen.size(), cudaMemcpyDeviceToHost, e->st));
        CK(cudaStreamSynchronize(e->st));
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        bool finite = true;
        for (int t = 0; t < n; t++)
            for (int c = 0; c < 3; c++) {
                const double v = cen[4 * (size_t)t + c];
                if (!(v > -1e15 && v < 1e15)) finite = false;
                lo[c] = std::min(lo[c], v); hi[c] = std::max(hi[c], v);
            }
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
    std::vector<long long> cnt(n), off(n + 1, 0);
    CK(cudaMemcpyAsync(cnt.data(), e->d_row_count, sizeof(long long) * n, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    for (int i = 0; i < n; i++) off[i + 1] = off[i] + cnt[i];
    const long long np = off[n];
    if (np >= (1LL << 31)) return fail(EDMD_ERR_LIMIT, "%lld pairs exceeds the 2^31 pair-index limit", np);
    if (int rc = ensure_pair_capacity(e, np)) return rc;
    CK(cudaMemcpyAsync(e->d_row_off, off.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, e->st));
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
    timer_end(e, T_PAIRS, cells ? 6 : 3);
    std::vector<int2> pairs((size_t)np);
    if (np) CK(cudaMemcpyAsync(pairs.data(), e->d_pairs, sizeof(int2) * np, cudaMemcpyDeviceToHost, e->st));
    CK(cudaStreamSynchronize(e->st));
    e->np = np;
    if (int rc = build_adjacency(e, pairs)) return rc;
    e->have_pairs = true; e->hits_valid = false; drop_graph(e);
    if (!e->shard_pairs_custom) { e->p0 = 0; e->p1 = np; }
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
    if (int rc = build_adjacency(e, pv)) return rc;
    e->have_pairs = true; e->hits_valid = false; drop_graph(e);
    if (!e->shard_pairs_custom) { e->p0 = 0; e->p1 = np; }
    return 0;
}

int edmd_set_shard(edmd_engine* e, int t0, int t1, long long p0, long long p1) {
    if (int rc = check_ready(e, false)) return rc;
    if (t0 < 0 || t1 > e->n || t0 > t1) return fail(EDMD_ERR_ARG, "tetrad shard [%d,%d) out of range", t0, t1);
    if (p0 < 0 || p0 > p1 || (e->have_pairs && p1 > e->np)) return fail(EDMD_ERR_ARG, "pair shard out of range");
    CK(cudaSetDevice(e->device));
    if (int rc = normalize(e)) return rc;
    const bool range_changed = (t0 != e->t0 || t1 != e->t1);
    drop_graph(e);
    e->t0 = t0; e->t1 = t1; e->p0 = p0; e->p1 = p1; e->shard_pairs_custom = true;
    e->hits_valid = false;
    if (range_changed) return upload_order(e);
    return 0;
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

static int one_step(edmd_engine* e, int random_mode, const long* calls_dev) {
    if (int rc = do_ed(e)) return rc;
    if (int rc = do_random(e, random_mode, e->time0, e->rank, e->calls, calls_dev)) return rc;
    if (random_mode == 1 && calls_dev) { CK(launch_bump_counter(e->d_calls, (long)e->n, e->st)); e->launches++; }
    if (int rc = do_scan(e)) return rc;
    return do_finish(e);
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
    const bool graph_ok = e->use_graph && !e->timing && e->nb_clip && !e->peer_mode && nsteps >= 1;
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
        CK(cudaGraphInstantiate(&e->graph_exec, e->graph, 0));
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
    if (random_mode != 0 && random_mode != 1) return fail(EDMD_ERR_ARG, "random mode %d unknown", random_mode);
    const bool pipelined = dense_field(e, tets, 0) && dense_field(e, tets, 1) && e->nb_clip && !e->peer_mode && !e->timing;
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

int edmd_sync(edmd_engine* e) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->st));
    return 0;
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

int edmd_nb_stats_ex(edmd_engine* e, long long out[4]) {
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
    out[0] = out[1] = out[2] = out[3] = 0;
    for (long long p = 0; p < e->np; p++) {
        if (!h[p]) continue;
        out[0]++;
        out[2] += __builtin_popcountll(h[p]);
        marked[pr[p].x] = 1; marked[pr[p].y] = 1;
    }
    for (int t = e->t0; t < e->t1; t++) out[1] += marked[t];
    if (e->scan_mode == 4 && e->d_pmask) {
        unsigned cnt = 0;
        CK(cudaMemcpy(&cnt, e->d_pmask, sizeof(unsigned), cudaMemcpyDeviceToHost));
        out[3] = cnt;
    }
    return 0;
}

void* edmd_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { fail(EDMD_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes); return nullptr; }
    return p;
}
void edmd_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int edmd_peer_hits_export(edmd_engine* e, unsigned char* handles /* 2 x 64 bytes */) {
    if (int rc = check_ready(e, true)) return rc;
    if (!handles) return fail(EDMD_ERR_ARG, "null handle buffer");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->st));
    close_peers(e);
    const long long need = std::max<long long>(e->np, 1);
    if (need > e->hitx_cap) {
        cudaFree(e->d_hitx[0]); cudaFree(e->d_hitx[1]); e->d_hitx[0] = e->d_hitx[1] = nullptr; e->hitx_cap = 0;
        const long long cap = need + need / 2 + 4096;
        CK(cudaMalloc((void**)&e->d_hitx[0], sizeof(unsigned long long) * (size_t)cap)); CK(cudaMalloc((void**)&e->d_hitx[1], sizeof(unsigned long long) * (size_t)cap));
        e->hitx_cap = cap;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    for (int b = 0; b < 2; b++) {
        CK(cudaMemset(e->d_hitx[b], 0, sizeof(unsigned long long) * (size_t)e->hitx_cap));
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, e->d_hitx[b]));
        memcpy(handles + 64 * b, &h, 64);
    }
    return 0;
}

int edmd_peer_hits_open(edmd_engine* e, const unsigned char* all_handles /* world x 2 x 64 */, int world, int rank) {
    if (int rc = check_ready(e, true)) return rc;
    if (!all_handles || world < 1 || world > 32 || rank < 0 || rank >= world || !e->d_hitx[0])
        return fail(EDMD_ERR_ARG, "edmd_peer_hits_open: bad arguments (export first; world <= 32)");
    CK(cudaSetDevice(e->device));
    close_peers(e);
    std::vector<unsigned long long*> tbl(2 * (size_t)world, nullptr);
    for (int r = 0; r < world; r++)
        for (int b = 0; b < 2; b++) {
            if (r == rank) { tbl[(size_t)b * world + r] = e->d_hitx[b]; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + ((size_t)r * 2 + b) * 64, 64);
            void* p = nullptr;
            CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            e->peer_opened.push_back(p);
            tbl[(size_t)b * world + r] = (unsigned long long*)p;
        }
    CK(cudaMalloc((void**)&e->d_peer_tbl, sizeof(unsigned long long*) * tbl.size()));
    CK(cudaMemcpy(e->d_peer_tbl, tbl.data(), sizeof(unsigned long long*) * tbl.size(), cudaMemcpyHostToDevice));
    e->world = world; e->rank_id = rank; e->hit_parity = 0; e->peer_mode = true;
    drop_graph(e);
    return 0;
}

int edmd_peer_hits_close(edmd_engine* e) {
    if (!e) return fail(EDMD_ERR_ARG, "null engine");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->st));
    close_peers(e);
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

}  // extern "C"
