// host_capi.cpp — C entry points over the product's C++ host layer (class IO / EDMD / Tetrad in
// msc-project_b200/host) so that pytest can drive it.  Mirrors oracle/ref_io_driver.cpp name for
// name (hostio_* vs refio_*); the shim_* calls exercise class EDMD method by method.
#include "io.hpp"

#include <unistd.h>
#include <cstring>

static IO* g_io = 0;
static EDMD* g_edmd = 0;

static Tetrad make_tetrad(int na, int nev, const double* avg, const double* masses, const double* abq,
                          const double* evals, const double* evecs, const double* crd, const double* vel) {
    Tetrad T; T.num_Atoms = na; T.num_Evecs = nev; T.allocate_Tetrad_Arrays();
    const int n3 = 3 * na;
    memcpy(T.avg, avg, 8 * n3); memcpy(T.masses, masses, 8 * n3); memcpy(T.abq, abq, 8 * n3);
    memcpy(T.eigenvalues, evals, 8 * nev); memcpy(T.eigenvectors[0], evecs, 8 * (size_t)n3 * nev);
    if (crd) memcpy(T.coordinates, crd, 8 * n3);
    if (vel) memcpy(T.velocities, vel, 8 * n3);
    return T;
}

extern "C" {

int hostio_load(const char* dir) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd))) return -1;
    if (chdir(dir) != 0) return -2;
    g_io = new IO();
    g_edmd = new EDMD();
    g_io->read_Cofig(g_edmd);
    g_io->read_Prm();
    g_io->read_Crd();
    g_io->initialise_Tetrad_Crds();
    if (chdir(cwd) != 0) return -3;
    return 0;
}
void hostio_counts(int* out) {
    out[0] = g_io->prm.num_Tetrads; out[1] = g_io->crd.num_BP; out[2] = g_io->crd.total_Atoms;
    out[3] = g_io->irest; out[4] = g_io->nsteps; out[5] = g_io->ntsync; out[6] = g_io->ntwt; out[7] = g_io->ntpr;
}
void hostio_params(double* p) {
    p[0] = g_edmd->dt; p[1] = g_edmd->gamma; p[2] = g_edmd->tautp; p[3] = g_edmd->temperature;
    p[4] = g_edmd->scaled; p[5] = g_edmd->mole_Cutoff; p[6] = g_edmd->atom_Cutoff; p[7] = g_edmd->mole_Least;
}
void hostio_tetrad_sizes(int t, int* na, int* nev) { *na = g_io->tetrad[t].num_Atoms; *nev = g_io->tetrad[t].num_Evecs; }
void hostio_tetrad(int t, double* avg, double* masses, double* abq, double* evals, double* evecs, double* crd, double* vel) {
    const Tetrad& T = g_io->tetrad[t];
    const int n3 = 3 * T.num_Atoms;
    memcpy(avg, T.avg, 8 * n3); memcpy(masses, T.masses, 8 * n3); memcpy(abq, T.abq, 8 * n3);
    memcpy(evals, T.eigenvalues, 8 * T.num_Evecs); memcpy(evecs, T.eigenvectors[0], 8 * (size_t)n3 * T.num_Evecs);
    memcpy(crd, T.coordinates, 8 * n3); memcpy(vel, T.velocities, 8 * n3);
}
void hostio_write(const char* dir, int istep, const double* energies, int index, const double* coordinates,
                  const double* velocities) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd)) || chdir(dir) != 0) return;
    double e[4] = {energies[0], energies[1], energies[2], energies[3]};
    g_io->write_Energies(istep + g_io->ntsync, e);
    g_io->write_Trajectory(istep, index, const_cast<double*>(coordinates));
    g_io->update_Crd_File(const_cast<double*>(velocities), const_cast<double*>(coordinates));
    if (chdir(cwd) != 0) return;
}
int hostio_displ(int b) { return g_io->displs[b]; }

// ---- class EDMD, one method per call (GPU) ------------------------------------------------------
static EDMD* shim(const double* p) {
    static EDMD e;
    e.initialise(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
    return &e;
}
void shim_ed(const double* p, int na, int nev, const double* avg, const double* masses, const double* abq,
             const double* evals, const double* evecs, double* crd_inout, double* ed_out /* 3N+1 */) {
    Tetrad T = make_tetrad(na, nev, avg, masses, abq, evals, evecs, crd_inout, 0);
    shim(p)->calculate_ED_Forces(&T);
    memcpy(crd_inout, T.coordinates, 8 * 3 * na); memcpy(ed_out, T.ED_Forces, 8 * (3 * na + 1));
    T.deallocate_Tetrad_Arrays();
}
long shim_random_calls(void) { return EDMD::random_calls; }
void shim_random(const double* p, int na, int nev, const double* avg, const double* masses, const double* abq,
                 const double* evals, const double* evecs, int rank, long pinned_time, double* out) {
    Tetrad T = make_tetrad(na, nev, avg, masses, abq, evals, evecs, 0, 0);
    EDMD* e = shim(p); e->pinned_time = pinned_time;
    e->calculate_Random_Terms(&T, rank);
    memcpy(out, T.random_Terms, 8 * 3 * na);
    T.deallocate_Tetrad_Arrays();
}
void shim_nb(const double* p, int na, int nev, const double* avg, const double* masses, const double* abq,
             const double* evals, const double* evecs, const double* crd1, const double* crd2,
             double* nb1 /* 3N+2 */, double* nb2) {
    Tetrad A = make_tetrad(na, nev, avg, masses, abq, evals, evecs, crd1, 0);
    Tetrad B = make_tetrad(na, nev, avg, masses, abq, evals, evecs, crd2, 0);
    shim(p)->calculate_NB_Forces(&A, &B);
    memcpy(nb1, A.NB_Forces, 8 * (3 * na + 2)); memcpy(nb2, B.NB_Forces, 8 * (3 * na + 2));
    A.deallocate_Tetrad_Arrays(); B.deallocate_Tetrad_Arrays();
}
void shim_integrate(const double* p, int na, int nev, const double* avg, const double* masses, const double* abq,
                    const double* evals, const double* evecs, double* crd_inout, double* vel_inout,
                    const double* ed, const double* rnd, const double* nb, double* temperature) {
    Tetrad T = make_tetrad(na, nev, avg, masses, abq, evals, evecs, crd_inout, vel_inout);
    memcpy(T.ED_Forces, ed, 8 * 3 * na); memcpy(T.random_Terms, rnd, 8 * 3 * na); memcpy(T.NB_Forces, nb, 8 * 3 * na);
    EDMD* e = shim(p);
    e->update_Velocities(&T);
    e->update_Coordinates(&T);
    memcpy(crd_inout, T.coordinates, 8 * 3 * na); memcpy(vel_inout, T.velocities, 8 * 3 * na);
    *temperature = T.temperature;
    T.deallocate_Tetrad_Arrays();
}

}
