/* edmd_b200.h — C ABI of the B200-native EDMD force-and-integrate engine.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has no FFI
 * layer: its seam is the C++ class EDMD (5 void methods on Tetrad*, /root/reference
 * src/edmd.hpp:54-170) driven by Worker::force_Calculation (src/worker.cpp:147-187) and
 * Master::update_Velocity/update_Coordinate/merge_Vels_n_Crds/generate_Pair_Lists
 * (src/master.cpp:157-384).  Each entry point below names the reference interface it
 * replaces.  Signatures use only plain pointers and sizes.
 *
 * All calls on one engine must come from one host thread at a time (the reference's
 * calculate_Random_Terms is not re-entrant either, edmd.cpp:173,179).  The engine copies
 * what it is given; it never keeps or frees caller memory.
 *
 * Every function returns 0 on success and a negative edmd_status on failure;
 * edmd_last_error() returns a description of the last failure on the calling thread.
 * There is NO CPU fallback: without a CUDA device edmd_create fails with EDMD_ERR_CUDA.
 */
#ifndef EDMD_B200_H
#define EDMD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDMD_B200_ABI_VERSION 2

typedef enum {
    EDMD_OK = 0,
    EDMD_ERR_ARG = -1,      /* bad argument / call order                              */
    EDMD_ERR_CUDA = -2,     /* CUDA runtime failure (no device, OOM, launch error)    */
    EDMD_ERR_STATE = -3,    /* engine not ready for this call (e.g. no pair list yet) */
    EDMD_ERR_LIMIT = -4     /* input exceeds a documented limit                       */
} edmd_status;

/* Simulation parameters in INTERNAL units — the eight public fields of class EDMD
 * (edmd.hpp:60-74) as EDMD::initialise receives them (edmd.cpp:51-60), i.e. after the unit
 * conversion of IO::read_Cofig (io.cpp:85-91): dt*=20.455, gamma/=20.455, tautp*=20.455,
 * scaled = 0.002*temperature. */
typedef struct edmd_params {
    double dt, gamma, tautp, temperature, scaled, mole_Cutoff, atom_Cutoff, mole_Least;
} edmd_params;

/* Byte-for-byte the data members of the reference's class Tetrad (tetrad.hpp:36-60) on
 * LP64: a `Tetrad*` from reference code can be passed as `edmd_tetrad*` unchanged.
 * eigenvectors is the row-pointer table of Array::allocate_2D_Double_Array
 * (array.cpp:24-34): eigenvectors[0] is a dense num_Evecs x 3*num_Atoms row-major block. */
typedef struct edmd_tetrad {
    int      num_Atoms;
    int      num_Evecs;
    double*  avg;           /* [3N] reference structure, xyz interleaved                */
    double*  masses;        /* [3N] each atom's mass replicated x3 (io.cpp:149-152)     */
    double*  abq;           /* [3N] (A,B,q) per atom; only q = abq[3i+2] is ever read   */
    double*  eigenvalues;   /* [num_Evecs]                                              */
    double** eigenvectors;  /* [num_Evecs][3N]                                          */
    double*  ED_Forces;     /* [3N+1] last = ED energy                                  */
    double*  random_Terms;  /* [3N]                                                     */
    double*  NB_Forces;     /* [3N+2] last two = NB energy, electrostatic energy        */
    double   temperature;
    double*  velocities;    /* [3N]                                                     */
    double*  coordinates;   /* [3N]                                                     */
} edmd_tetrad;

typedef struct edmd_engine edmd_engine;

const char* edmd_last_error(void);
int         edmd_abi_version(void);

/* ---- lifetime ---------------------------------------------------------------------- */
/* Replaces EDMD::EDMD + EDMD::initialise (edmd.cpp:23-60).  `device` is the CUDA ordinal. */
int  edmd_create(const edmd_params* p, int device, edmd_engine** out);
void edmd_destroy(edmd_engine* e);
int  edmd_set_params(edmd_engine* e, const edmd_params* p);
/* krep / qfac are literals in the reference (edmd.cpp:237,242: 100.0, 0.0).  Exposed so
 * that electrostatics can be switched on (qfac = 332.064/4 per the comment at :240). */
int  edmd_set_nb_consts(edmd_engine* e, double krep, double qfac);

/* The clip of summed NB force components to [-1,1] belongs to Master::process_NB_Forces
 * (master.cpp:304-305), not to EDMD::calculate_NB_Forces; on by default, switched off by the
 * class-EDMD shim to return raw pair forces. */
int  edmd_set_nb_clip(edmd_engine* e, int on);
/* Arithmetic form of the NB distance pass (nb_scan_kernel): 0 = direct differences exactly
 * as edmd.cpp:251-256, 1 = pair-local Gram form |xi|^2+|xj|^2-2xi.xj (3 instead of 6 FP64
 * instructions per atom pair), 2 = the same test in FP32 with packed FFMA2 and a proven error
 * margin (the "FP32 variant" of BASELINE configs[4]), 4 = form 1 plus exact bounding-sphere
 * culling of whole tetrad pairs and of 32 x 32 atom blocks (default).  Modes 0-2 evaluate all
 * 63,504 atom pairs of every listed tetrad pair like the reference does; mode 4 skips blocks that
 * provably hold no atom pair inside the cutoff.  The pass only selects tetrad pairs for the exact
 * FP64 force kernel, so forces and energies are bit-identical in all modes. */
int  edmd_set_nb_scan_mode(edmd_engine* e, int mode);
/* Projection kernel of the ED pass (edmd.cpp:106-150): 0 = automatic (default), 1 = one CTA per tetrad
 * with the eigenvector coefficients in registers (any number of eigenvectors), 2 = one warp per
 * tetrad with the parameter set staged in shared memory (sets with <= 12 eigenvectors; chosen
 * automatically when tetrads share parameter sets), 3 = the two passes as FP64 tensor-core contractions
 * (DMMA.8x8x4) over groups of eight tetrads that share a set (sets with <= 12 eigenvectors; falls back to 1
 * otherwise).  Same arithmetic up to the order of the projection sums. */
int  edmd_set_ed_kernel(edmd_engine* e, int which);

/* ---- data in / out ----------------------------------------------------------------- */
/* Replaces Master::send_Tetrads / Worker::recv_Tetrads (master.cpp:137-153, MPI_Tetrad
 * datatype mpilib.cpp:23-62): uploads avg, masses, abq, eigenvalues, eigenvectors of n
 * tetrads plus their coordinates and velocities.  Identical parameter sets (same bytes)
 * are stored once on the device.  bp_atoms (n+3 entries, the per-base-pair atom counts of
 * the .crd file, io.cpp:201-205) is needed only by edmd_merge and may be NULL. */
int edmd_upload_tetrads(edmd_engine* e, const edmd_tetrad* tets, int n, const int* bp_atoms);
/* coordinates + velocities host -> device / device -> host (MPI_Crds, mpilib.cpp:106-130) */
int edmd_set_state(edmd_engine* e, const edmd_tetrad* tets, int n);
int edmd_get_state(edmd_engine* e, edmd_tetrad* tets, int n);
/* The same for this engine's owned tetrads only (edmd_set_shard): what one rank of a multi-GPU run moves per
 * step.  Needs the coordinates (velocities) of all tetrads in one host block each (edmd_alloc_pinned).
 * edmd_set_state_owned is asynchronous (stream-ordered before the next step), edmd_get_state_owned returns
 * when the owned block of the host arrays is current. */
int edmd_set_state_owned(edmd_engine* e, const edmd_tetrad* tets, int n);
int edmd_get_state_owned(edmd_engine* e, edmd_tetrad* tets, int n);
/* Fills ED_Forces[3N+1], random_Terms[3N], NB_Forces[3N+2] and temperature exactly as the
 * reference lays them out (MPI_ED_Forces, mpilib.cpp:74-94; master.cpp:308,316-317). */
int edmd_get_forces(edmd_engine* e, edmd_tetrad* tets, int n);
/* Overwrites the device force arrays (used to test update_Velocities in isolation). */
int edmd_set_forces(edmd_engine* e, const edmd_tetrad* tets, int n);

/* Page-locked host memory for the caller's Tetrad arrays.  When the coordinates (velocities)
 * of all tetrads lie back to back in one block, edmd_set_state/edmd_get_state copy straight
 * between that block and the device; if the block is pinned the copy runs at full PCIe rate. */
void* edmd_alloc_pinned(size_t bytes);
void  edmd_free_pinned(void* p);

/* Flat variants for callers that do not hold Tetrad structs: arrays are tetrad-
 * concatenated xyz-interleaved, tetrad t at 3*sum_{s<t} num_Atoms[s].  NULL = skip. */
int edmd_set_state_flat(edmd_engine* e, const double* crd, const double* vel);
int edmd_get_state_flat(edmd_engine* e, double* crd, double* vel);
int edmd_get_forces_flat(edmd_engine* e, double* ed, double* ed_energy, double* rnd,
                         double* nb, double* nb_energy /* 2 per tetrad */, double* temperature);

/* ---- pair list --------------------------------------------------------------------- */
/* Replaces Master::cal_Centre_of_Mass + generate_Pair_Lists (master.cpp:157-210): same
 * inclusion predicate, same (t1,t2) orientation, same order; 64-bit count. */
int edmd_build_pairlist(edmd_engine* e, long long* num_pairs);
/* 0 = auto, 1 = the reference's O(n^2) sweep, 2 = uniform-grid cell list (same list, same order). */
int edmd_set_pair_builder(edmd_engine* e, int which);
int edmd_get_pairlist(edmd_engine* e, int* pairs /* 2*num_pairs */, long long cap_pairs);
/* Install a caller-made list (tests, custom schedules). */
int edmd_set_pairlist(edmd_engine* e, const int* pairs, long long num_pairs);
/* Multi-GPU: restrict this engine's NB distance pass to pairs [p0,p1) of the list and its
 * per-tetrad work (ED, accumulate, integrate, merge output) to tetrads [t0,t1).
 * Defaults: everything. */
int edmd_set_shard(edmd_engine* e, int t0, int t1, long long p0, long long p1);

/* ---- multi-GPU: one engine per GPU, exchanging through peer memory over NVLink/NVSwitch -------------
 * Replaces the reference's per-step messages (master.cpp:268-292, worker.cpp:109-142; datatypes
 * mpilib.cpp:23-130).  Every rank holds the whole system (as every MPI worker did, worker.cpp:67-87) and
 * owns a contiguous block of ceil(n/world) tetrads plus a slice of the pair list (edmd_set_shard).  With
 * peers attached, edmd_step runs the sharded step — ED and random terms for owned tetrads; flag barrier;
 * pull of the non-owned tetrads this rank reads, straight from their owners' memory (MPI_Bcast(MPI_Crds),
 * master.cpp:279); distance pass over the pair slice, whose kernel stores each flagged pair's block mask
 * into EVERY rank's buffer (instead of the dense MPI_Reduce, master.cpp:288); flag barrier; force pass +
 * clip + integrate for owned tetrads — all kernels on the engine stream, replayed as one CUDA graph.
 * No force is ever reduced across ranks, so results are bit-identical for every world size.
 *
 * One process per GPU (torchrun): after the pair list exists every rank calls edmd_peer_export, the ranks
 * exchange the blobs (any transport), call edmd_peer_attach with all of them in rank order, then
 * edmd_set_shard.  When a later edmd_build_pairlist fails with EDMD_ERR_STATE because the list outgrew
 * the mask buffer: edmd_peer_detach everywhere, export and attach again.  One process driving several
 * GPUs: the edmd_group_* calls below do all of this internally. */
#define EDMD_PEER_BLOB_BYTES 320
int edmd_peer_export(edmd_engine* e, unsigned char* blob /* EDMD_PEER_BLOB_BYTES */);
int edmd_peer_attach(edmd_engine* e, const unsigned char* blobs /* world * EDMD_PEER_BLOB_BYTES */, int world, int rank);
int edmd_peer_detach(edmd_engine* e);
/* Collective: make this rank's copy of every rank's owned data complete.  what = 1 coordinates | 2 velocities
 * | 4 per-tetrad energies and temperatures.  Needed before edmd_merge, edmd_build_pairlist, edmd_get_state
 * and edmd_get_energies; every rank must call it (it contains two barriers). */
int edmd_peer_gather(edmd_engine* e, int what);
int edmd_world(edmd_engine* e, int* world, int* rank);

/* Single process, several GPUs (the C++ driver): a group owns one engine per device, keeps them in step
 * and exposes the same operations as a single engine.  n_gpus engines on `devices` (NULL = 0..n_gpus-1).
 * This is the `edmd_create(params, n_gpus, ...)` of a multi-GPU caller. */
typedef struct edmd_group edmd_group;
int  edmd_group_create(const edmd_params* p, int n_gpus, const int* devices, edmd_group** out);
void edmd_group_destroy(edmd_group* g);
int  edmd_group_size(edmd_group* g);
edmd_engine* edmd_group_engine(edmd_group* g, int rank);   /* for the per-engine setters */
int  edmd_group_upload_tetrads(edmd_group* g, const edmd_tetrad* tets, int n, const int* bp_atoms);
int  edmd_group_set_state(edmd_group* g, const edmd_tetrad* tets, int n);
int  edmd_group_get_state(edmd_group* g, edmd_tetrad* tets, int n);
int  edmd_group_set_rng(edmd_group* g, long time0, int rank, long calls_before);
/* edmd_build_pairlist on every GPU (the list is a pure function of the centroids) + the work split:
 * contiguous tetrad blocks, equal pair slices (replaces Master::generate_Indexes, master.cpp:214-242). */
int  edmd_group_build_pairlist(edmd_group* g, long long* num_pairs);
int  edmd_group_step(edmd_group* g, int nsteps, int random_mode);
int  edmd_group_merge(edmd_group* g);
int  edmd_group_get_energies(edmd_group* g, double out[4]);
int  edmd_group_snapshot_begin(edmd_group* g, int ticket, double* crd, double* vel, double* obs);
int  edmd_group_snapshot_wait(edmd_group* g, int ticket);
int  edmd_group_sync(edmd_group* g);

/* ---- the five EDMD methods, batched over all (owned) tetrads ---------------------------- */
/* EDMD::calculate_ED_Forces (edmd.cpp:64-166) incl. CalcRMSDRotationalMatrix
 * (qcprot.c:392-407).  Overwrites coordinates with the re-embedded ones, like the reference. */
int edmd_ed_forces(edmd_engine* e);
/* EDMD::calculate_Random_Terms (edmd.cpp:170-228).  mode 0: zero terms ("random off").
 * mode 1: the reference stream — glibc TYPE_3 rand() seeded per call with
 * (unsigned)(13579 + calls + rank^3 + time0), call counter advancing one per tetrad with
 * the wrap of edmd.cpp:180, Kinderman-Monahan ratio of uniforms (ACM 712). */
int edmd_random_terms(edmd_engine* e, int mode, long time0, int rank, long calls_before);
/* Worker::force_Calculation NB section + Master::process_NB_Forces (worker.cpp:165-179,
 * master.cpp:297-322): EDMD::calculate_NB_Forces (edmd.cpp:232-285) over the pair list,
 * per-tetrad sums in pair-list order, then clip of force components to [-1,1]. */
int edmd_nb_forces(edmd_engine* e);
/* The two halves of edmd_nb_forces, for multi-GPU drivers: distance pass over this
 * engine's pair shard (fills the per-pair hit flags), then accumulate + clip on owners. */
int edmd_nb_scan(edmd_engine* e);
int edmd_nb_accumulate(edmd_engine* e);
/* EDMD::update_Velocities (edmd.cpp:289-317) and EDMD::update_Coordinates (:322-327). */
int edmd_update_velocities(edmd_engine* e);
int edmd_update_coordinates(edmd_engine* e);
/* edmd_nb_accumulate + edmd_integrate fused (the clipped force never leaves registers
 * between the two); what edmd_step runs after the distance pass. */
int edmd_nb_finish(edmd_engine* e);
/* update_Velocity then update_Coordinate back to back (simulation.cpp:45-46) in one pass. */
int edmd_integrate(edmd_engine* e);
/* Master::merge_Vels_n_Crds (master.cpp:347-384). */
int edmd_merge(edmd_engine* e);
/* Master::write_Info sums (master.cpp:388-401): {sum E_ed, sum E_nb, sum E_el, mean T}. */
int edmd_get_energies(edmd_engine* e, double out[4]);

/* ---- fused step -------------------------------------------------------------------- */
/* nsteps x { calculate_Forces; update_Velocity; update_Coordinate } (simulation.cpp:42-48)
 * with W = 1 semantics.  random_mode as in edmd_random_terms; the call counter continues
 * from the value passed to edmd_set_rng. */
int edmd_set_rng(edmd_engine* e, long time0, int rank, long calls_before);
int edmd_step(edmd_engine* e, int nsteps, int random_mode);
/* edmd_set_state + edmd_step + edmd_get_state in one call, in place on the caller's Tetrad
 * array (Master's view of a step: host tetrads in, host tetrads out, master.cpp:268-343).  When
 * the coordinate and velocity arrays are one slab each (edmd_alloc_pinned), the velocity upload
 * overlaps the ED and distance kernels.  Same results as the three separate calls. */
int edmd_step_host(edmd_engine* e, edmd_tetrad* tets, int n, int nsteps, int random_mode);
/* Asynchronous output on a side stream (replaces the blocking gathers of Master::write_Info / write_Crds,
 * master.cpp:388-419): the coordinates and velocities as they are at this point of the run (reference
 * layout, tetrads concatenated, xyz interleaved: edmd_state_doubles() doubles each) and the per-tetrad
 * observables obs[4n] = { E_ed[n] | (E_nb, E_el)[n] | temperature[n] } drain into the caller's (pinned)
 * buffers while the engine goes on stepping.  Two tickets (0, 1) may be in flight; edmd_snapshot_wait blocks
 * until a ticket's buffers are complete (callable from a writer thread).  NULL pointers are skipped. */
int edmd_snapshot_begin(edmd_engine* e, int ticket, double* crd, double* vel, double* obs);
int edmd_snapshot_wait(edmd_engine* e, int ticket);
long long edmd_state_doubles(edmd_engine* e);
/* Optional: run the random-term generator on a parallel branch of the step graph beside the ED and distance
 * kernels (0 = off, the default: measured slower on B200; 1 = on; 2..16 = on with that many
 * generator CTAs resident per SM).  Same results either way. */
int edmd_set_rng_branch(edmd_engine* e, int on);
/* edmd_step replays a CUDA graph of one step (on by default; off: plain kernel launches). */
int edmd_set_graph(edmd_engine* e, int on);
int edmd_sync(edmd_engine* e);

/* ---- introspection (bench / multi-GPU plumbing) -------------------------------------- */
typedef enum {
    EDMD_BUF_CRD = 0,      /* double [n][3][S]   coordinates, SoA planes per tetrad */
    EDMD_BUF_VEL = 1,
    EDMD_BUF_FED = 2,
    EDMD_BUF_RND = 3,
    EDMD_BUF_FNB = 4,
    EDMD_BUF_HIT = 5,      /* unsigned char [num_pairs] per-pair hit flags           */
    EDMD_BUF_CENTROID = 6  /* double [n][4]                                          */
} edmd_buffer;
/* Device pointer + byte size of an engine buffer (for NCCL collectives issued by the
 * host-side driver through torch.distributed).  Valid until the next upload. */
int edmd_device_buffer(edmd_engine* e, int which, void** dptr, size_t* bytes);
int edmd_atom_stride(edmd_engine* e);      /* S: atoms per tetrad plane, multiple of 32   */
int edmd_num_tetrads(edmd_engine* e);
long long edmd_num_pairs(edmd_engine* e);
void* edmd_stream(edmd_engine* e);         /* the cudaStream_t all engine work runs on    */
/* Device time of everything enqueued between the two calls, CUDA events on the engine
 * stream (torch.cuda.Event would only see torch's current stream).  stop synchronises. */
int edmd_timer_start(edmd_engine* e);
int edmd_timer_stop(edmd_engine* e, double* ms);
/* Non-blocking variant: record numbered marks (CUDA events) on the engine stream and read the
 * device time between two of them later (edmd_mark_elapsed synchronises on slot_b). */
int edmd_mark(edmd_engine* e, int slot);
int edmd_mark_elapsed(edmd_engine* e, int slot_a, int slot_b, double* ms);
/* Accumulated device time (ms) and launch count per kernel since the last reset, measured
 * with CUDA events on the engine stream.  which: 0 ED, 1 NB scan, 2 NB accumulate+clip,
 * 3 random, 4 merge, 5 pair list, 6 integrate, 7 layout (pack/unpack at the boundary), 8 the superposition
 * kernel alone (part of 0). */
int edmd_timing_enable(edmd_engine* e, int on);
int edmd_timing_get(edmd_engine* e, int which, double* total_ms, long long* launches);
int edmd_timing_reset(edmd_engine* e);
long long edmd_launch_count(edmd_engine* e);   /* kernels launched since create */
/* Atom pairs inside atom_Cutoff / tetrad pairs flagged, from the last NB pass. */
int edmd_nb_stats(edmd_engine* e, long long* flagged_pairs);
/* More of the same: out = {tetrad pairs flagged, owned tetrads in at least one flagged pair, 32 x 32 atom
 * blocks the force pass evaluates (set bits of the per-pair block masks), tetrad pairs the bounding-sphere
 * tests of the culled distance pass let through, 32 x 32 atom blocks the culled distance kernel evaluated
 * (set bits of the sphere masks); the last two are 0 in the brute-force modes}. */
int edmd_nb_stats_ex(edmd_engine* e, long long out[5]);

/* FP64 peak of this device measured with a dependent-free DFMA loop (the roofline
 * denominator for the NB kernel; MEASURED_PEAKS.json has no FP64 entry). */
int edmd_measure_fp64_peak(int device, double* tflops, double* sm_mhz_est);

#ifdef __cplusplus
}
#endif
#endif /* EDMD_B200_H */
