/* ref_driver.cpp — drives the UNMODIFIED reference physics.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiled together with /root/reference/src/{edmd,tetrad,array}.cpp and
 * /root/reference/src/qcprot/qcprot.c (as C++) into oracle/_ref/libedmd_ref.so by
 * oracle/Makefile.  Every number this library produces for the five EDMD methods comes
 * from the reference's own object code; this file only restates the step protocol that
 * lives in the MPI-dependent files, which cannot be built here (no MPI in the image):
 *
 *   pair list          master.cpp:157-210   (Master::cal_Centre_of_Mass, generate_Pair_Lists)
 *   block split        master.cpp:214-242   (Master::generate_Indexes)
 *   force evaluation   worker.cpp:147-199   (Worker::force_Calculation, empty_NB_Forces)
 *   clip               master.cpp:297-322   (Master::process_NB_Forces)
 *   integrate          master.cpp:327-343, simulation.cpp:42-48
 *   merge              master.cpp:347-384   (Master::merge_Vels_n_Crds)
 *   energies           master.cpp:388-401   (Master::write_Info)
 *
 * Deviations from the reference protocol, all deliberate and documented in DESIGN.md:
 *   - W = 1 worker ordering: every NB interaction sees post-shake coordinates.
 *   - the NB accumulator is zero-filled before use (reference: uninitialised on the master).
 *   - pair indices are int, the pair count is 64-bit (reference: doubles / int overflow).
 *   - "random_on = 0" sets random_Terms to zero (the reference has no such switch).
 *   - time() is pinned (orc_set_time) so calculate_Random_Terms is reproducible.
 */
#include "edmd.hpp"      /* /root/reference/src/edmd.hpp (pulls tetrad.hpp, array.hpp, qcprot.h) */
#include "oracle_api.h"

#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

/* ---- time() pin: bound inside this .so by -Wl,-Bsymbolic-functions ------------------ */
static long g_time0 = 1000000000L;
static long g_rng_calls = 0;
extern "C" time_t time(time_t* out) {
    if (out) *out = (time_t)g_time0;
    return (time_t)g_time0;
}

struct orc_sys {
    int n;
    EDMD edmd;
    Tetrad* tet;
    std::vector<long long> off3;   /* n+1 */
    int max_atoms;
    std::vector<int> pairs;        /* 2*np */
    long long np;
};

extern "C" {

const char* orc_kind(void) { return "reference"; }

orc_sys* orc_create(int n, const int* num_atoms, const int* num_evecs) {
    orc_sys* s = new orc_sys();
    s->n = n;
    s->tet = new Tetrad[n];
    s->off3.assign(n + 1, 0);
    s->max_atoms = 0;
    s->np = 0;
    for (int t = 0; t < n; t++) {
        s->tet[t].num_Atoms = num_atoms[t];
        s->tet[t].num_Evecs = num_evecs[t];
        s->tet[t].allocate_Tetrad_Arrays();               /* tetrad.cpp:23-36 */
        s->off3[t + 1] = s->off3[t] + 3LL * num_atoms[t];
        if (num_atoms[t] > s->max_atoms) s->max_atoms = num_atoms[t];
        const int n3 = 3 * num_atoms[t];
        /* io.cpp:276-285 initial values */
        memset(s->tet[t].velocities, 0, sizeof(double) * n3);
        memset(s->tet[t].coordinates, 0, sizeof(double) * n3);
        memset(s->tet[t].ED_Forces, 0, sizeof(double) * (n3 + 1));
        memset(s->tet[t].random_Terms, 0, sizeof(double) * n3);
        memset(s->tet[t].NB_Forces, 0, sizeof(double) * (n3 + 2));
        s->tet[t].temperature = 0.0;
    }
    return s;
}

void orc_destroy(orc_sys* s) {
    if (!s) return;
    for (int t = 0; t < s->n; t++) s->tet[t].deallocate_Tetrad_Arrays();
    delete[] s->tet;
    delete s;
}

void orc_set_params(orc_sys* s, const double p[8]) {
    s->edmd.initialise(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
}

int orc_set_nb_consts(orc_sys*, double krep, double qfac) {
    return (krep == 100.0 && qfac == 0.0) ? 0 : -1;
}

void orc_set_static(orc_sys* s, int t, const double* avg, const double* masses3,
                    const double* abq, const double* evals, const double* evecs) {
    Tetrad& T = s->tet[t];
    const int n3 = 3 * T.num_Atoms;
    memcpy(T.avg, avg, sizeof(double) * n3);
    memcpy(T.masses, masses3, sizeof(double) * n3);
    memcpy(T.abq, abq, sizeof(double) * n3);
    memcpy(T.eigenvalues, evals, sizeof(double) * T.num_Evecs);
    memcpy(T.eigenvectors[0], evecs, sizeof(double) * (size_t)T.num_Evecs * n3);
}

void orc_set_state(orc_sys* s, const double* crd, const double* vel) {
    for (int t = 0; t < s->n; t++) {
        const int n3 = 3 * s->tet[t].num_Atoms;
        if (crd) memcpy(s->tet[t].coordinates, crd + s->off3[t], sizeof(double) * n3);
        if (vel) memcpy(s->tet[t].velocities, vel + s->off3[t], sizeof(double) * n3);
    }
}

void orc_get_state(orc_sys* s, double* crd, double* vel) {
    for (int t = 0; t < s->n; t++) {
        const int n3 = 3 * s->tet[t].num_Atoms;
        if (crd) memcpy(crd + s->off3[t], s->tet[t].coordinates, sizeof(double) * n3);
        if (vel) memcpy(vel + s->off3[t], s->tet[t].velocities, sizeof(double) * n3);
    }
}

void orc_get_forces(orc_sys* s, double* ed, double* ed_energy, double* rnd, double* nb,
                    double* nb_energy, double* temperature) {
    for (int t = 0; t < s->n; t++) {
        const Tetrad& T = s->tet[t];
        const int n3 = 3 * T.num_Atoms;
        if (ed) memcpy(ed + s->off3[t], T.ED_Forces, sizeof(double) * n3);
        if (ed_energy) ed_energy[t] = T.ED_Forces[n3];
        if (rnd) memcpy(rnd + s->off3[t], T.random_Terms, sizeof(double) * n3);
        if (nb) memcpy(nb + s->off3[t], T.NB_Forces, sizeof(double) * n3);
        if (nb_energy) { nb_energy[2 * t] = T.NB_Forces[n3]; nb_energy[2 * t + 1] = T.NB_Forces[n3 + 1]; }
        if (temperature) temperature[t] = T.temperature;
    }
}

void orc_set_forces(orc_sys* s, const double* ed, const double* rnd, const double* nb) {
    for (int t = 0; t < s->n; t++) {
        Tetrad& T = s->tet[t];
        const int n3 = 3 * T.num_Atoms;
        if (ed) memcpy(T.ED_Forces, ed + s->off3[t], sizeof(double) * n3);
        if (rnd) memcpy(T.random_Terms, rnd + s->off3[t], sizeof(double) * n3);
        if (nb) memcpy(T.NB_Forces, nb + s->off3[t], sizeof(double) * n3);
    }
}

void orc_ed_forces(orc_sys* s, int t) { s->edmd.calculate_ED_Forces(&s->tet[t]); }
void orc_random_terms(orc_sys* s, int t, int rank) {
    s->edmd.calculate_Random_Terms(&s->tet[t], rank);
    g_rng_calls++;
}
void orc_nb_forces(orc_sys* s, int t1, int t2) { s->edmd.calculate_NB_Forces(&s->tet[t1], &s->tet[t2]); }
void orc_update_velocities(orc_sys* s, int t) { s->edmd.update_Velocities(&s->tet[t]); }
void orc_update_coordinates(orc_sys* s, int t) { s->edmd.update_Coordinates(&s->tet[t]); }

double orc_qcp(const double* ref_xyz, const double* x_xyz, int n, double rot[9]) {
    /* same transposition the caller does at edmd.cpp:77-85 */
    double** a = Array::allocate_2D_Double_Array(3, n);
    double** b = Array::allocate_2D_Double_Array(3, n);
    for (int i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) { a[c][i] = ref_xyz[3 * i + c]; b[c][i] = x_xyz[3 * i + c]; }
    double r = CalcRMSDRotationalMatrix(a, b, n, rot, NULL);
    Array::deallocate_2D_Double_Array(a);
    Array::deallocate_2D_Double_Array(b);
    return r;
}

void orc_set_time(long t0) { g_time0 = t0; }
long orc_rng_calls(void) { return g_rng_calls; }
int  orc_set_rng_calls(long) { return -1; }

/* master.cpp:157-210 */
long long orc_build_pairs(orc_sys* s) {
    const int n = s->n;
    std::vector<double> com(3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        double cx = 0.0, cy = 0.0, cz = 0.0;
        const Tetrad& T = s->tet[i];
        for (int j = 0; j < 3 * T.num_Atoms; j += 3) {
            cx += T.coordinates[j]; cy += T.coordinates[j + 1]; cz += T.coordinates[j + 2];
        }
        com[3 * i] = cx / T.num_Atoms; com[3 * i + 1] = cy / T.num_Atoms; com[3 * i + 2] = cz / T.num_Atoms;
    }
    s->pairs.clear();
    const double cut2 = s->edmd.mole_Cutoff * s->edmd.mole_Cutoff;
    for (int i = 0; i < n; i++) {
        for (int j = i + 1; j < n; j++) {
            const double r = (com[3*i] - com[3*j]) * (com[3*i] - com[3*j])
                           + (com[3*i+1] - com[3*j+1]) * (com[3*i+1] - com[3*j+1])
                           + (com[3*i+2] - com[3*j+2]) * (com[3*i+2] - com[3*j+2]);
            const int sep = j - i;   /* abs(i - j), j > i */
            if (r < cut2 && sep > s->edmd.mole_Least && sep < (n - s->edmd.mole_Least)) {
                /* master.cpp:197-201: odd separation -> (j,i), even -> (i,j) */
                if (sep % 2 == 1) { s->pairs.push_back(j); s->pairs.push_back(i); }
                else              { s->pairs.push_back(i); s->pairs.push_back(j); }
            }
        }
    }
    s->np = (long long)(s->pairs.size() / 2);
    return s->np;
}

long long orc_num_pairs(orc_sys* s) { return s->np; }
void orc_get_pairs(orc_sys* s, int* out) { if (s->np) memcpy(out, s->pairs.data(), sizeof(int) * 2 * s->np); }
void orc_set_pairs(orc_sys* s, long long np, const int* pairs) {
    s->pairs.assign(pairs, pairs + 2 * np);
    s->np = np;
}

/* contiguous near-equal split, master.cpp:214-242 */
static void block_split(long long total, int workers, int w, long long* lo, long long* hi) {
    long long base = total / workers, rem = total - base * workers;
    *lo = w * base + (w < rem ? w : rem);
    *hi = *lo + base + (w < rem ? 1 : 0);
}

/* NB part of worker.cpp:165-179 for pairs [p0,p1) into acc[n][stride]; the worker's private
 * replica of each tetrad's NB_Forces scratch is a shallow Tetrad copy with its own buffer. */
static void nb_block(orc_sys* s, long long p0, long long p1, double* acc, int stride,
                     double* scratch1, double* scratch2) {
    for (long long p = p0; p < p1; p++) {
        const int i1 = s->pairs[2 * p], i2 = s->pairs[2 * p + 1];
        Tetrad a = s->tet[i1], b = s->tet[i2];   /* shallow: pointers alias the shared state */
        a.NB_Forces = scratch1; b.NB_Forces = scratch2;
        s->edmd.calculate_NB_Forces(&a, &b);
        double* r1 = acc + (size_t)i1 * stride;
        double* r2 = acc + (size_t)i2 * stride;
        for (int j = 0; j < 3 * a.num_Atoms + 2; j++) r1[j] += scratch1[j];
        for (int j = 0; j < 3 * b.num_Atoms + 2; j++) r2[j] += scratch2[j];
    }
}

void orc_forces(orc_sys* s, int random_on, int nthreads) {
    const int n = s->n;
    const int stride = 3 * s->max_atoms + 2;
    if (nthreads < 1) nthreads = 1;

    /* ED + random, worker.cpp:155-162 (rank 1 = the single worker) */
    if (nthreads == 1) {
        for (int t = 0; t < n; t++) {
            s->edmd.calculate_ED_Forces(&s->tet[t]);
            if (random_on) { s->edmd.calculate_Random_Terms(&s->tet[t], 1); g_rng_calls++; }
            else memset(s->tet[t].random_Terms, 0, sizeof(double) * 3 * s->tet[t].num_Atoms);
        }
    } else {
        /* ED in parallel blocks; random terms stay serial: libc rand() is one global stream */
        std::vector<std::thread> th;
        for (int w = 0; w < nthreads; w++)
            th.emplace_back([s, n, nthreads, w]() {
                long long lo, hi; block_split(n, nthreads, w, &lo, &hi);
                for (long long t = lo; t < hi; t++) s->edmd.calculate_ED_Forces(&s->tet[t]);
            });
        for (auto& x : th) x.join();
        for (int t = 0; t < n; t++) {
            if (random_on) { s->edmd.calculate_Random_Terms(&s->tet[t], 1); g_rng_calls++; }
            else memset(s->tet[t].random_Terms, 0, sizeof(double) * 3 * s->tet[t].num_Atoms);
        }
    }

    /* NB: zeroed accumulator (worker.cpp:191-199), pairs in list order */
    std::vector<double> acc((size_t)n * stride, 0.0);
    if (nthreads == 1) {
        std::vector<double> s1(stride), s2(stride);
        nb_block(s, 0, s->np, acc.data(), stride, s1.data(), s2.data());
    } else {
        std::vector<std::vector<double> > part(nthreads);
        std::vector<std::thread> th;
        for (int w = 0; w < nthreads; w++)
            th.emplace_back([s, n, stride, nthreads, w, &part]() {
                long long lo, hi; block_split(s->np, nthreads, w, &lo, &hi);
                part[w].assign((size_t)n * stride, 0.0);
                std::vector<double> s1(stride), s2(stride);
                nb_block(s, lo, hi, part[w].data(), stride, s1.data(), s2.data());
            });
        for (auto& x : th) x.join();
        /* MPI_Reduce(MPI_SUM), worker.cpp:182 / master.cpp:288 */
        for (int w = 0; w < nthreads; w++)
            for (size_t k = 0; k < acc.size(); k++) acc[k] += part[w][k];
    }

    /* clip, master.cpp:297-322 */
    for (int t = 0; t < n; t++) {
        Tetrad& T = s->tet[t];
        const int n3 = 3 * T.num_Atoms;
        double* r = acc.data() + (size_t)t * stride;
        for (int j = 0; j < n3; j++) {
            if (r[j] < -1.0) r[j] = -1.0; else if (r[j] > 1.0) r[j] = 1.0;
            T.NB_Forces[j] = r[j];
        }
        T.NB_Forces[n3] = r[n3];
        T.NB_Forces[n3 + 1] = r[n3 + 1];
    }
}

void orc_step(orc_sys* s, int nsteps, int random_on, int nthreads) {
    for (int it = 0; it < nsteps; it++) {
        orc_forces(s, random_on, nthreads);
        for (int t = 0; t < s->n; t++) s->edmd.update_Velocities(&s->tet[t]);   /* master.cpp:327-333 */
        for (int t = 0; t < s->n; t++) s->edmd.update_Coordinates(&s->tet[t]);  /* master.cpp:337-343 */
    }
}

double orc_time_nb(orc_sys* s, long long p0, long long p1, int nthreads) {
    const int n = s->n;
    const int stride = 3 * s->max_atoms + 2;
    if (nthreads < 1) nthreads = 1;
    if (p1 > s->np) p1 = s->np;
    std::vector<std::vector<double> > part(nthreads);
    for (int w = 0; w < nthreads; w++) part[w].assign((size_t)n * stride, 0.0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int w = 0; w < nthreads; w++)
        th.emplace_back([s, stride, nthreads, w, p0, p1, &part]() {
            long long lo, hi; block_split(p1 - p0, nthreads, w, &lo, &hi);
            std::vector<double> s1(stride), s2(stride);
            nb_block(s, p0 + lo, p0 + hi, part[w].data(), stride, s1.data(), s2.data());
        });
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

/* master.cpp:347-384 with io.cpp:236-247 displacements */
void orc_merge(orc_sys* s, const int* bp_atoms) {
    const int n = s->n, nbp = n + 3;
    std::vector<long long> displs(nbp + 1, 0);
    for (int i = 1; i <= nbp; i++) displs[i] = displs[i - 1] + 3LL * bp_atoms[i - 1];
    const long long tot = displs[nbp];
    std::vector<double> vel(tot, 0.0), crd(tot, 0.0);
    for (int i = 0; i < n; i++) {
        long long index = displs[i];
        for (int j = 0; j < 3 * s->tet[i].num_Atoms; j++, index++) {
            vel[index] += s->tet[i].velocities[j];
            crd[index] += s->tet[i].coordinates[j];
        }
    }
    long long index = displs[nbp - 3];
    for (long long i = 0; index < displs[nbp]; index++, i++) {
        vel[index] += vel[i]; vel[i] = vel[index];
        crd[index] += crd[i]; crd[i] = crd[index];
    }
    for (long long i = 0; i < tot; i++) { vel[i] *= 0.25; crd[i] *= 0.25; }
    for (int i = 0; i < n; i++) {
        long long idx = displs[i];
        for (int j = 0; j < 3 * s->tet[i].num_Atoms; j++, idx++) {
            s->tet[i].velocities[j] = vel[idx];
            s->tet[i].coordinates[j] = crd[idx];
        }
    }
}

void orc_energies(orc_sys* s, double e[4]) {
    e[0] = e[1] = e[2] = e[3] = 0.0;
    for (int i = 0; i < s->n; i++) {
        const Tetrad& T = s->tet[i];
        e[0] += T.ED_Forces[3 * T.num_Atoms];
        e[1] += T.NB_Forces[3 * T.num_Atoms];
        e[2] += T.NB_Forces[3 * T.num_Atoms + 1];
        e[3] += T.temperature;
    }
    e[3] /= s->n;
}

} /* extern "C" */
