/* ref_io_driver.cpp — C entry points over the UNMODIFIED reference IO class.  TEST INFRASTRUCTURE ONLY.
 * Built by oracle/Makefile (target refio) from /root/reference/src/io.cpp (+ edmd/tetrad/array/qcprot)
 * into oracle/_ref/libedmd_refio.so.  The reference reads "./config.txt" from the current directory
 * (io.cpp:67), so refio_load chdir()s into the deck directory first. */
#include "io.hpp"   /* /root/reference/src/io.hpp */

#include <unistd.h>
#include <cstring>

static IO* g_io = 0;
static EDMD* g_edmd = 0;

extern "C" {

int refio_load(const char* dir) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd))) return -1;
    if (chdir(dir) != 0) return -2;
    g_io = new IO();          /* leaked on purpose: the reference's ~IO never frees the tetrad array itself */
    g_edmd = new EDMD();
    g_io->read_Cofig(g_edmd);
    g_io->read_Prm();
    g_io->read_Crd();
    g_io->initialise_Tetrad_Crds();
    if (chdir(cwd) != 0) return -3;
    return 0;
}
void refio_counts(int* out /* num_Tetrads, num_BP, total_Atoms, irest, nsteps, ntsync, ntwt, ntpr */) {
    out[0] = g_io->prm.num_Tetrads; out[1] = g_io->crd.num_BP; out[2] = g_io->crd.total_Atoms;
    out[3] = g_io->irest; out[4] = g_io->nsteps; out[5] = g_io->ntsync; out[6] = g_io->ntwt; out[7] = g_io->ntpr;
}
void refio_params(double* p /* 8 */) {
    p[0] = g_edmd->dt; p[1] = g_edmd->gamma; p[2] = g_edmd->tautp; p[3] = g_edmd->temperature;
    p[4] = g_edmd->scaled; p[5] = g_edmd->mole_Cutoff; p[6] = g_edmd->atom_Cutoff; p[7] = g_edmd->mole_Least;
}
void refio_tetrad_sizes(int t, int* na, int* nev) { *na = g_io->tetrad[t].num_Atoms; *nev = g_io->tetrad[t].num_Evecs; }
void refio_tetrad(int t, double* avg, double* masses, double* abq, double* evals, double* evecs, double* crd, double* vel) {
    const Tetrad& T = g_io->tetrad[t];
    const int n3 = 3 * T.num_Atoms;
    memcpy(avg, T.avg, 8 * n3); memcpy(masses, T.masses, 8 * n3); memcpy(abq, T.abq, 8 * n3);
    memcpy(evals, T.eigenvalues, 8 * T.num_Evecs); memcpy(evecs, T.eigenvectors[0], 8 * (size_t)n3 * T.num_Evecs);
    memcpy(crd, T.coordinates, 8 * n3); memcpy(vel, T.velocities, 8 * n3);
}
/* writers: paths are the ones read from config.txt, relative to `dir` */
void refio_write(const char* dir, int istep, const double* energies, int index, const double* coordinates,
                 const double* velocities) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd)) || chdir(dir) != 0) return;
    double e[4] = {energies[0], energies[1], energies[2], energies[3]};
    g_io->write_Energies(istep + g_io->ntsync, e);
    g_io->write_Trajectory(istep, index, const_cast<double*>(coordinates));
    g_io->update_Crd_File(const_cast<double*>(velocities), const_cast<double*>(coordinates));
    if (chdir(cwd) != 0) return;
}
int refio_displ(int b) { return g_io->displs[b]; }

}
