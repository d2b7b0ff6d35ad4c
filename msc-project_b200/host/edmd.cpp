// edmd.cpp — class EDMD as batch-of-one calls into libedmd_b200 (see edmd.hpp).
#include "edmd.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

extern "C" {
#include "../../include/edmd_b200.h"
}

static_assert(sizeof(Tetrad) == sizeof(edmd_tetrad), "Tetrad must match edmd_tetrad");

namespace {
void die_on(int rc, const char* what) {
    if (rc == 0) return;
    // the reference's methods are void; its fatal paths print and exit(1) (io.cpp:107-108)
    std::fprintf(stderr, ">>> ERROR: %s failed: %s\n", what, edmd_last_error());
    std::exit(1);
}
edmd_tetrad* as_c(Tetrad* t) { return reinterpret_cast<edmd_tetrad*>(t); }
}  // namespace

EDMD::EDMD(void) {
    // defaults of the reference constructor (edmd.cpp:23-43): 2 fs, 2 /ps, 0.2 ps, 300 K
    constants.Boltzmann = 0.002;
    constants.timefac = 20.455;
    dt = 0.002 * constants.timefac;
    gamma = 2.0 / constants.timefac;
    tautp = 0.2 * constants.timefac;
    temperature = 300.0;
    scaled = constants.Boltzmann * temperature;
    mole_Cutoff = 30.0;
    atom_Cutoff = 10.0;
    mole_Least = 5.0;
    device = 0;
    pinned_time = -1;
    one = two = 0;
}

EDMD::~EDMD(void) {
    if (one) edmd_destroy(one);
    if (two) edmd_destroy(two);
}

void EDMD::initialise(double _dt, double _gamma, double _tautp, double _temperature, double _scaled,
                      double _mole_Cutoff, double _atom_Cutoff, double _mole_Least) {
    dt = _dt; gamma = _gamma; tautp = _tautp; temperature = _temperature; scaled = _scaled;
    mole_Cutoff = _mole_Cutoff; atom_Cutoff = _atom_Cutoff; mole_Least = _mole_Least;
}

void EDMD::sync_params(edmd_engine* e) {
    edmd_params p = {dt, gamma, tautp, temperature, scaled, mole_Cutoff, atom_Cutoff, mole_Least};
    die_on(edmd_set_params(e, &p), "edmd_set_params");
}

edmd_engine* EDMD::engine_for(Tetrad* t) {
    if (!one) die_on(edmd_create(0, device, &one), "edmd_create");
    sync_params(one);
    // Always a full upload: the caller may have changed any array in place since the last call
    // (the reference's methods read the Tetrad afresh every time).
    die_on(edmd_upload_tetrads(one, as_c(t), 1, 0), "edmd_upload_tetrads");
    return one;
}

edmd_engine* EDMD::engine_for(Tetrad* t1, Tetrad* t2) {
    if (!two) die_on(edmd_create(0, device, &two), "edmd_create");
    sync_params(two);
    Tetrad both[2] = {*t1, *t2};   // shallow copies: the arrays stay the caller's
    die_on(edmd_upload_tetrads(two, as_c(both), 2, 0), "edmd_upload_tetrads");
    const int pair[2] = {0, 1};
    die_on(edmd_set_pairlist(two, pair, 1), "edmd_set_pairlist");
    die_on(edmd_set_nb_clip(two, 0), "edmd_set_nb_clip");   // the clip belongs to Master::process_NB_Forces
    return two;
}

// edmd.cpp:64-166 — fills ED_Forces[0..3N] (last = energy) and overwrites coordinates
void EDMD::calculate_ED_Forces(Tetrad* tetrad) {
    edmd_engine* e = engine_for(tetrad);
    die_on(edmd_ed_forces(e), "edmd_ed_forces");
    const int n3 = 3 * tetrad->num_Atoms;
    die_on(edmd_get_forces_flat(e, tetrad->ED_Forces, tetrad->ED_Forces + n3, 0, 0, 0, 0), "edmd_get_forces_flat");
    die_on(edmd_get_state_flat(e, tetrad->coordinates, 0), "edmd_get_state_flat");
}

// edmd.cpp:170-228 — the seed counter is per process like the reference's function-static
long EDMD::random_calls = 0;

void EDMD::calculate_Random_Terms(Tetrad* tetrad, int rank) {
    edmd_engine* e = engine_for(tetrad);
    const long t0 = pinned_time >= 0 ? pinned_time : (long)std::time(0);
    die_on(edmd_random_terms(e, 1, t0, rank, random_calls), "edmd_random_terms");
    random_calls++;
    die_on(edmd_get_forces_flat(e, 0, 0, tetrad->random_Terms, 0, 0, 0), "edmd_get_forces_flat");
}

// edmd.cpp:232-285 — unclipped pair forces on both tetrads, pair energies credited to both
void EDMD::calculate_NB_Forces(Tetrad* t1, Tetrad* t2) {
    edmd_engine* e = engine_for(t1, t2);
    die_on(edmd_nb_forces(e), "edmd_nb_forces");
    const int n1 = 3 * t1->num_Atoms, n2 = 3 * t2->num_Atoms;
    std::vector<double> f(n1 + n2);
    double en[4];
    die_on(edmd_get_forces_flat(e, 0, 0, 0, f.data(), en, 0), "edmd_get_forces_flat");
    std::memcpy(t1->NB_Forces, f.data(), sizeof(double) * n1);
    std::memcpy(t2->NB_Forces, f.data() + n1, sizeof(double) * n2);
    t1->NB_Forces[n1] = en[0]; t1->NB_Forces[n1 + 1] = en[1];
    t2->NB_Forces[n2] = en[0]; t2->NB_Forces[n2 + 1] = en[1];   // edmd.cpp:282-283: copied from t1
}

// edmd.cpp:289-317
void EDMD::update_Velocities(Tetrad* tetrad) {
    edmd_engine* e = engine_for(tetrad);
    die_on(edmd_set_forces(e, as_c(tetrad), 1), "edmd_set_forces");
    die_on(edmd_update_velocities(e), "edmd_update_velocities");
    die_on(edmd_get_state_flat(e, 0, tetrad->velocities), "edmd_get_state_flat");
    die_on(edmd_get_forces_flat(e, 0, 0, 0, 0, 0, &tetrad->temperature), "edmd_get_forces_flat");
}

// edmd.cpp:322-327
void EDMD::update_Coordinates(Tetrad* tetrad) {
    edmd_engine* e = engine_for(tetrad);
    die_on(edmd_update_coordinates(e), "edmd_update_coordinates");
    die_on(edmd_get_state_flat(e, tetrad->coordinates, 0), "edmd_get_state_flat");
}
