/* oracle_api.h — C interface shared by the two CPU checkers.  TEST INFRASTRUCTURE ONLY.
 *
 * Two shared objects export exactly these symbols:
 *
 *   oracle/_ref/libedmd_ref.so   "reference": the UNMODIFIED reference physics
 *       (/root/reference/src/edmd.cpp, tetrad.cpp, array.cpp, qcprot/qcprot.c compiled
 *       where they lie) driven by oracle/ref_driver.cpp, which restates only the
 *       master/worker step protocol (MPI is not available in this image).
 *   oracle/_build/libedmd_oracle.so   "port": oracle/edmd_oracle.c, a plain-C restatement
 *       of the same arithmetic (travels to the GPU box as source + built .so).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load either library.  The product (msc-project_b200/) never does.
 *
 * Flat layout used by every bulk call: tetrads are concatenated; tetrad t owns
 * doubles [off3[t], off3[t] + 3*num_atoms[t]) of a "3N" array, where
 * off3[t] = 3 * sum_{s<t} num_atoms[s].  Per-tetrad arrays are xyz-interleaved exactly
 * as the reference's Tetrad fields (tetrad.hpp:36-60).
 */
#ifndef EDMD_ORACLE_API_H
#define EDMD_ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_sys orc_sys;

/* "reference" or "port" */
const char* orc_kind(void);

orc_sys* orc_create(int n_tetrads, const int* num_atoms, const int* num_evecs);
void     orc_destroy(orc_sys* s);

/* p = {dt, gamma, tautp, temperature, scaled, mole_Cutoff, atom_Cutoff, mole_Least}
 * in INTERNAL units, i.e. after the conversion of io.cpp:85-91 (EDMD::initialise,
 * edmd.cpp:51-60). */
void orc_set_params(orc_sys* s, const double p[8]);

/* krep / qfac are hard-coded 100.0 / 0.0 in the reference (edmd.cpp:237,242).  The port
 * accepts other values (returns 0); the reference build returns -1 unless they match. */
int  orc_set_nb_consts(orc_sys* s, double krep, double qfac);

/* static per-tetrad data: avg[3N], masses[3N] (each mass already replicated x3,
 * io.cpp:149-152), abq[3N], eigenvalues[nev], eigenvectors[nev][3N] row-major. */
void orc_set_static(orc_sys* s, int t, const double* avg, const double* masses3,
                    const double* abq, const double* evals, const double* evecs);

/* bulk dynamic state, flat layout */
void orc_set_state(orc_sys* s, const double* crd, const double* vel);
void orc_get_state(orc_sys* s, double* crd, double* vel);
/* ed: flat 3N per tetrad; ed_energy[n]; rnd flat; nb flat; nb_energy[2n] (NB, electrostatic);
 * temperature[n].  Any pointer may be NULL. */
void orc_get_forces(orc_sys* s, double* ed, double* ed_energy, double* rnd, double* nb,
                    double* nb_energy, double* temperature);
void orc_set_forces(orc_sys* s, const double* ed, const double* rnd, const double* nb);

/* The five EDMD methods, one tetrad (pair) at a time — edmd.cpp:64,170,232,289,322. */
void orc_ed_forces(orc_sys* s, int t);
void orc_random_terms(orc_sys* s, int t, int rank);
void orc_nb_forces(orc_sys* s, int t1, int t2);
void orc_update_velocities(orc_sys* s, int t);
void orc_update_coordinates(orc_sys* s, int t);

/* QCP alone (qcprot.c:392-407): ref/x are xyz-interleaved [3n]; rot is row-major 3x3
 * rotating x onto ref.  Returns the RMSD. */
double orc_qcp(const double* ref_xyz, const double* x_xyz, int n, double rot[9]);

/* RNG pinning.  calculate_Random_Terms seeds libc rand() with
 * (unsigned)(RNG_Seed++ + rank^3 + time(NULL)) (edmd.cpp:173,179-180).  orc_set_time pins
 * time(); orc_rng_calls returns how many times calculate_Random_Terms ran in this
 * process (the function-static RNG_Seed cannot be reset in the reference build). */
void orc_set_time(long t0);
long orc_rng_calls(void);
/* port only (reference build: no-op returning -1): set the call counter. */
int  orc_set_rng_calls(long calls);

/* Pair list, master.cpp:157-210.  Returns the pair count (64-bit). */
long long orc_build_pairs(orc_sys* s);
long long orc_num_pairs(orc_sys* s);
void orc_get_pairs(orc_sys* s, int* pairs /* 2*np: t1,t2 in list order */);
void orc_set_pairs(orc_sys* s, long long np, const int* pairs);

/* One force evaluation = Master::calculate_Forces with W = 1 worker semantics
 * (worker.cpp:147-187 then master.cpp:297-322): ED for every tetrad (coordinates are
 * overwritten by the shaken ones), random terms if random_on (else zero), NB over the
 * pair list accumulated in list order, then clip to [-1,1].
 * nthreads > 1 maps the reference's contiguous block split (master.cpp:214-242) onto
 * host threads — used only for CPU baseline timing (summation order then differs). */
void orc_forces(orc_sys* s, int random_on, int nthreads);
/* nsteps x { orc_forces; update_Velocities; update_Coordinates } (simulation.cpp:42-48) */
void orc_step(orc_sys* s, int nsteps, int random_on, int nthreads);
/* Only the NB part over pairs [p0, p1): returns seconds spent (baseline sampling). */
double orc_time_nb(orc_sys* s, long long p0, long long p1, int nthreads);

/* 4-fold overlap average, master.cpp:347-384.  Requires equal atoms per base pair
 * bookkeeping: bp_atoms[n_tetrads+3]. */
void orc_merge(orc_sys* s, const int* bp_atoms);
/* {sum E_ed, sum E_nb, sum E_el, mean T}  master.cpp:388-401 */
void orc_energies(orc_sys* s, double e[4]);

#ifdef __cplusplus
}
#endif
#endif
