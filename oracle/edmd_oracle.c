/* edmd_oracle.c — plain-C CPU restatement of the reference's EDMD hot path ("port").
 *
 * TEST INFRASTRUCTURE ONLY.  Loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; never by the product (msc-project_b200/).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit against
 * oracle/_ref/libedmd_ref.so (the unmodified reference sources compiled where they lie)
 * whenever /root/reference is present, and against tests/golden/*.npz vectors generated
 * from that same build (tests/golden/make_golden.py) everywhere else.  The reference ships
 * no golden vectors of its own; its two source-comment known answers
 * (edmd.cpp:182 and :298) are checked in tests/test_oracle.py as well.
 *
 * Every function cites the reference lines it follows.  Expression trees are kept
 * identical to the reference's so that, compiled with -ffp-contract=off on x86-64, the
 * results are bitwise equal to the reference build (which uses no FMA either).
 *
 * Third-party pieces restated here because they are not under /root/reference:
 *   glibc rand()/srand() (TYPE_3 additive feedback, glibc 2.39 stdlib/random_r.c) —
 *   call sites edmd.cpp:179,206,207.  Checked against this container's libc in
 *   tests/test_oracle.py.
 */
#include "oracle_api.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#define KB_INTERNAL 0.002 /* Constants.Boltzmann, edmd.cpp:25 */

struct orc_sys {
    int n;
    int* na;            /* atoms per tetrad */
    int* nev;           /* eigenvectors per tetrad */
    long long* off3;    /* n+1 offsets into 3N arrays */
    long long* offev;   /* n+1 offsets into eigenvalue array */
    long long* offevec; /* n+1 offsets into eigenvector array */
    int max_atoms;
    /* params, edmd.hpp:60-74 */
    double dt, gamma, tautp, temperature, scaled, mole_cutoff, atom_cutoff, mole_least;
    double krep, qfac;
    /* static */
    double *avg, *mass, *abq, *eval, *evec;
    /* dynamic */
    double *crd, *vel, *fed, *rnd, *fnb;
    double *e_ed, *e_nb /* 2 per tetrad */, *temp;
    /* pairs */
    int* pairs;
    long long np, pairs_cap;
};

static long g_time0 = 1000000000L;
static long g_rng_calls = 0;

const char* orc_kind(void) { return "port"; }

orc_sys* orc_create(int n, const int* num_atoms, const int* num_evecs) {
    orc_sys* s = (orc_sys*)calloc(1, sizeof(orc_sys));
    s->n = n;
    s->na = (int*)malloc(sizeof(int) * n);
    s->nev = (int*)malloc(sizeof(int) * n);
    s->off3 = (long long*)calloc(n + 1, sizeof(long long));
    s->offev = (long long*)calloc(n + 1, sizeof(long long));
    s->offevec = (long long*)calloc(n + 1, sizeof(long long));
    for (int t = 0; t < n; t++) {
        s->na[t] = num_atoms[t];
        s->nev[t] = num_evecs[t];
        s->off3[t + 1] = s->off3[t] + 3LL * num_atoms[t];
        s->offev[t + 1] = s->offev[t] + num_evecs[t];
        s->offevec[t + 1] = s->offevec[t] + 3LL * num_atoms[t] * num_evecs[t];
        if (num_atoms[t] > s->max_atoms) s->max_atoms = num_atoms[t];
    }
    const size_t n3 = (size_t)s->off3[n];
    s->avg = (double*)calloc(n3, 8); s->mass = (double*)calloc(n3, 8); s->abq = (double*)calloc(n3, 8);
    s->eval = (double*)calloc((size_t)s->offev[n] + 1, 8);
    s->evec = (double*)calloc((size_t)s->offevec[n] + 1, 8);
    s->crd = (double*)calloc(n3, 8); s->vel = (double*)calloc(n3, 8);
    s->fed = (double*)calloc(n3, 8); s->rnd = (double*)calloc(n3, 8); s->fnb = (double*)calloc(n3, 8);
    s->e_ed = (double*)calloc(n, 8); s->e_nb = (double*)calloc(2 * (size_t)n, 8); s->temp = (double*)calloc(n, 8);
    /* EDMD::EDMD defaults, edmd.cpp:23-43 */
    s->dt = 0.002 * 20.455; s->gamma = 2.0 / 20.455; s->tautp = 0.2 * 20.455;
    s->temperature = 300.0; s->scaled = KB_INTERNAL * 300.0;
    s->mole_cutoff = 30.0; s->atom_cutoff = 10.0; s->mole_least = 5.0;
    s->krep = 100.0; s->qfac = 0.0; /* edmd.cpp:237,242 */
    return s;
}

void orc_destroy(orc_sys* s) {
    if (!s) return;
    free(s->na); free(s->nev); free(s->off3); free(s->offev); free(s->offevec);
    free(s->avg); free(s->mass); free(s->abq); free(s->eval); free(s->evec);
    free(s->crd); free(s->vel); free(s->fed); free(s->rnd); free(s->fnb);
    free(s->e_ed); free(s->e_nb); free(s->temp); free(s->pairs);
    free(s);
}

/* EDMD::initialise, edmd.cpp:51-60 */
void orc_set_params(orc_sys* s, const double p[8]) {
    s->dt = p[0]; s->gamma = p[1]; s->tautp = p[2]; s->temperature = p[3];
    s->scaled = p[4]; s->mole_cutoff = p[5]; s->atom_cutoff = p[6]; s->mole_least = p[7];
}

int orc_set_nb_consts(orc_sys* s, double krep, double qfac) { s->krep = krep; s->qfac = qfac; return 0; }

void orc_set_static(orc_sys* s, int t, const double* avg, const double* masses3,
                    const double* abq, const double* evals, const double* evecs) {
    const size_t n3 = 3 * (size_t)s->na[t];
    memcpy(s->avg + s->off3[t], avg, 8 * n3);
    memcpy(s->mass + s->off3[t], masses3, 8 * n3);
    memcpy(s->abq + s->off3[t], abq, 8 * n3);
    memcpy(s->eval + s->offev[t], evals, 8 * (size_t)s->nev[t]);
    memcpy(s->evec + s->offevec[t], evecs, 8 * n3 * s->nev[t]);
}

void orc_set_state(orc_sys* s, const double* crd, const double* vel) {
    if (crd) memcpy(s->crd, crd, 8 * (size_t)s->off3[s->n]);
    if (vel) memcpy(s->vel, vel, 8 * (size_t)s->off3[s->n]);
}
void orc_get_state(orc_sys* s, double* crd, double* vel) {
    if (crd) memcpy(crd, s->crd, 8 * (size_t)s->off3[s->n]);
    if (vel) memcpy(vel, s->vel, 8 * (size_t)s->off3[s->n]);
}
void orc_get_forces(orc_sys* s, double* ed, double* ed_energy, double* rnd, double* nb,
                    double* nb_energy, double* temperature) {
    const size_t b = 8 * (size_t)s->off3[s->n];
    if (ed) memcpy(ed, s->fed, b);
    if (ed_energy) memcpy(ed_energy, s->e_ed, 8 * (size_t)s->n);
    if (rnd) memcpy(rnd, s->rnd, b);
    if (nb) memcpy(nb, s->fnb, b);
    if (nb_energy) memcpy(nb_energy, s->e_nb, 16 * (size_t)s->n);
    if (temperature) memcpy(temperature, s->temp, 8 * (size_t)s->n);
}
void orc_set_forces(orc_sys* s, const double* ed, const double* rnd, const double* nb) {
    const size_t b = 8 * (size_t)s->off3[s->n];
    if (ed) memcpy(s->fed, ed, b);
    if (rnd) memcpy(s->rnd, rnd, b);
    if (nb) memcpy(s->fnb, nb, b);
}

/* ------------------------------------------------------------------------------------
 * QCP superposition (Theobald 2005; Liu, Agrafiotis & Theobald 2010), as vendored in
 * qcprot.c v1.4.  x is rotated onto ref.  Inputs are 3 x n SoA and are centred in place
 * (qcprot.c:343-388 with weight == NULL), exactly like the reference.
 * ---------------------------------------------------------------------------------- */
static void centre_soa(double* x, double* y, double* z, int n) { /* qcprot.c:370-387 */
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = 0; i < n; i++) { sx += x[i]; sy += y[i]; sz += z[i]; }
    sx /= n; sy /= n; sz /= n;
    for (int i = 0; i < n; i++) { x[i] -= sx; y[i] -= sy; z[i] -= sz; }
}

/* qcprot.c:132-160: M = sum ref (x) mov, returns (G_ref + G_mov)/2 */
static double cross_covariance(double M[9], const double* rx, const double* ry, const double* rz,
                               const double* mx, const double* my, const double* mz, int n) {
    double g_ref = 0.0, g_mov = 0.0;
    for (int k = 0; k < 9; k++) M[k] = 0.0;
    for (int i = 0; i < n; i++) {
        const double a = rx[i], b = ry[i], c = rz[i];
        g_ref += a * a + b * b + c * c;
        const double p = mx[i], q = my[i], r = mz[i];
        g_mov += (p * p + q * q + r * r);
        M[0] += (a * p); M[1] += (a * q); M[2] += (a * r);
        M[3] += (b * p); M[4] += (b * q); M[5] += (b * r);
        M[6] += (c * p); M[7] += (c * q); M[8] += (c * r);
    }
    return (g_ref + g_mov) * 0.5;
}

/* qcprot.c:164-340 (minScore < 0 path).  Returns the RMSD. */
static double qcp_rotation(double R[9], const double M[9], double e0, int n) {
    const double xx = M[0], xy = M[1], xz = M[2];
    const double yx = M[3], yy = M[4], yz = M[5];
    const double zx = M[6], zy = M[7], zz = M[8];

    const double xx2 = xx * xx, yy2 = yy * yy, zz2 = zz * zz;
    const double xy2 = xy * xy, yz2 = yz * yz, xz2 = xz * xz;
    const double yx2 = yx * yx, zy2 = zy * zy, zx2 = zx * zx;

    /* characteristic polynomial lambda^4 + c2 lambda^2 + c1 lambda + c0, qcprot.c:202-223 */
    const double t_a = 2.0 * (yz * zy - yy * zz);
    const double t_b = yy2 + zz2 - xx2 + yz2 + zy2;
    const double c2 = -2.0 * (xx2 + yy2 + zz2 + xy2 + yx2 + xz2 + zx2 + yz2 + zy2);
    const double c1 = 8.0 * (xx * yz * zy + yy * zx * xz + zz * xy * yx - xx * yy * zz - yz * zx * xy - zy * yx * xz);

    const double xz_p = xz + zx, yz_p = yz + zy, xy_p = xy + yx;
    const double yz_m = yz - zy, xz_m = xz - zx, xy_m = xy - yx;
    const double xxyy_p = xx + yy, xxyy_m = xx - yy;
    const double t_c = xy2 + xz2 - yx2 - zx2;

    const double c0 = t_c * t_c
        + (t_b + t_a) * (t_b - t_a)
        + (-(xz_p) * (yz_m) + (xy_m) * (xxyy_m - zz)) * (-(xz_m) * (yz_p) + (xy_m) * (xxyy_m + zz))
        + (-(xz_p) * (yz_p) - (xy_p) * (xxyy_p - zz)) * (-(xz_m) * (yz_m) - (xy_p) * (xxyy_p + zz))
        + (+(xy_p) * (yz_p) + (xz_p) * (xxyy_m + zz)) * (-(xy_m) * (yz_m) + (xz_p) * (xxyy_p + zz))
        + (+(xy_p) * (yz_m) + (xz_m) * (xxyy_m - zz)) * (-(xy_m) * (yz_p) + (xz_m) * (xxyy_p - zz));

    /* Newton-Raphson from the upper bound e0, qcprot.c:226-238 */
    double lam = e0;
    for (int it = 0; it < 50; it++) {
        const double prev = lam;
        const double l2 = lam * lam;
        const double b = (l2 + c2) * lam;
        const double a = b + c1;
        const double delta = ((a * lam + c0) / (2.0 * l2 * lam + b + a));
        lam -= delta;
        if (fabs(lam - prev) < fabs(1e-11 * lam)) break;
    }
    const double rms = sqrt(fabs(2.0 * (e0 - lam) / n)); /* qcprot.c:244 */

    /* K - lambda I (symmetric 4x4), qcprot.c:252-255 */
    const double a11 = xxyy_p + zz - lam, a12 = yz_m, a13 = -xz_m, a14 = xy_m;
    const double a21 = yz_m, a22 = xxyy_m - zz - lam, a23 = xy_p, a24 = xz_p;
    const double a31 = a13, a32 = a23, a33 = yy - xx - zz - lam, a34 = yz_p;
    const double a41 = a14, a42 = a24, a43 = a34, a44 = zz - xxyy_p - lam;

    /* 2x2 minors of rows 3,4 — qcprot.c:256-258 */
    const double m3344 = a33 * a44 - a43 * a34, m3244 = a32 * a44 - a42 * a34;
    const double m3243 = a32 * a43 - a42 * a33, m3143 = a31 * a43 - a41 * a33;
    const double m3144 = a31 * a44 - a41 * a34, m3142 = a31 * a42 - a41 * a32;

    /* adjoint column 1, qcprot.c:259-264 */
    double q1 =  a22 * m3344 - a23 * m3244 + a24 * m3243;
    double q2 = -a21 * m3344 + a23 * m3144 - a24 * m3143;
    double q3 =  a21 * m3244 - a22 * m3144 + a24 * m3142;
    double q4 = -a21 * m3243 + a22 * m3143 - a23 * m3142;
    double qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;

    if (qq < 1e-6) { /* fallbacks, qcprot.c:271-309 */
        q1 =  a12 * m3344 - a13 * m3244 + a14 * m3243;
        q2 = -a11 * m3344 + a13 * m3144 - a14 * m3143;
        q3 =  a11 * m3244 - a12 * m3144 + a14 * m3142;
        q4 = -a11 * m3243 + a12 * m3143 - a13 * m3142;
        qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
        if (qq < 1e-6) {
            const double m1324 = a13 * a24 - a14 * a23, m1224 = a12 * a24 - a14 * a22;
            const double m1223 = a12 * a23 - a13 * a22, m1124 = a11 * a24 - a14 * a21;
            const double m1123 = a11 * a23 - a13 * a21, m1122 = a11 * a22 - a12 * a21;
            q1 =  a42 * m1324 - a43 * m1224 + a44 * m1223;
            q2 = -a41 * m1324 + a43 * m1124 - a44 * m1123;
            q3 =  a41 * m1224 - a42 * m1124 + a44 * m1122;
            q4 = -a41 * m1223 + a42 * m1123 - a43 * m1122;
            qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
            if (qq < 1e-6) {
                q1 =  a32 * m1324 - a33 * m1224 + a34 * m1223;
                q2 = -a31 * m1324 + a33 * m1124 - a34 * m1123;
                q3 =  a31 * m1224 - a32 * m1124 + a34 * m1122;
                q4 = -a31 * m1223 + a32 * m1123 - a33 * m1122;
                qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
                if (qq < 1e-6) { /* identity, qcprot.c:299-306 */
                    R[0] = R[4] = R[8] = 1.0;
                    R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0.0;
                    return rms;
                }
            }
        }
    }
    const double nq = sqrt(qq);
    q1 /= nq; q2 /= nq; q3 /= nq; q4 /= nq;
    /* unit quaternion -> rotation matrix, qcprot.c:317-337 */
    const double w2 = q1 * q1, i2 = q2 * q2, j2 = q3 * q3, k2 = q4 * q4;
    const double ij = q2 * q3, wk = q1 * q4, ki = q4 * q2, wj = q1 * q3, jk = q3 * q4, wi = q1 * q2;
    R[0] = w2 + i2 - j2 - k2; R[1] = 2 * (ij + wk);     R[2] = 2 * (ki - wj);
    R[3] = 2 * (ij - wk);     R[4] = w2 - i2 + j2 - k2; R[5] = 2 * (jk + wi);
    R[6] = 2 * (ki + wj);     R[7] = 2 * (jk - wi);     R[8] = w2 - i2 - j2 + k2;
    return rms;
}

/* qcprot.c:392-407 */
static double qcp_superpose(double* ref[3], double* mov[3], int n, double R[9]) {
    double M[9];
    centre_soa(ref[0], ref[1], ref[2], n);
    centre_soa(mov[0], mov[1], mov[2], n);
    const double e0 = cross_covariance(M, ref[0], ref[1], ref[2], mov[0], mov[1], mov[2], n);
    return qcp_rotation(R, M, e0, n);
}

double orc_qcp(const double* ref_xyz, const double* x_xyz, int n, double rot[9]) {
    double* buf = (double*)malloc(sizeof(double) * 6 * (size_t)n);
    double* a[3] = { buf, buf + n, buf + 2 * n };
    double* b[3] = { buf + 3 * n, buf + 4 * n, buf + 5 * n };
    for (int i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) { a[c][i] = ref_xyz[3 * i + c]; b[c][i] = x_xyz[3 * i + c]; }
    const double r = qcp_superpose(a, b, n, rot);
    free(buf);
    return r;
}

/* ------------------------------------------------------------------------------------
 * EDMD::calculate_ED_Forces, edmd.cpp:64-166
 * ---------------------------------------------------------------------------------- */
void orc_ed_forces(orc_sys* s, int t) {
    const int na = s->na[t], nev = s->nev[t], n3 = 3 * na;
    double* x = s->crd + s->off3[t];
    double* f = s->fed + s->off3[t];
    const double* avg = s->avg + s->off3[t];
    const double* ev = s->evec + s->offevec[t];
    const double* lam = s->eval + s->offev[t];
    double R[9], shift[3];

    double* work = (double*)malloc(sizeof(double) * (6 * (size_t)na + n3 + nev + n3));
    double* ac[3] = { work, work + na, work + 2 * na };                 /* centred avg  */
    double* xc[3] = { work + 3 * na, work + 4 * na, work + 5 * na };    /* centred x    */
    double* d = work + 6 * na;                                          /* 3N scratch   */
    double* proj = d + n3;
    double* frot = proj + nev;

    for (int i = 0; i < na; i++)                                        /* :77-85 */
        for (int c = 0; c < 3; c++) { ac[c][i] = avg[3 * i + c]; xc[c][i] = x[3 * i + c]; }

    qcp_superpose(ac, xc, na, R);                                       /* :88 */
    for (int i = 0; i < na; i++) {                                      /* :89-94 */
        d[3 * i]     = R[0] * xc[0][i] + R[1] * xc[1][i] + R[2] * xc[2][i] - ac[0][i];
        d[3 * i + 1] = R[3] * xc[0][i] + R[4] * xc[1][i] + R[5] * xc[2][i] - ac[1][i];
        d[3 * i + 2] = R[6] * xc[0][i] + R[7] * xc[1][i] + R[8] * xc[2][i] - ac[2][i];
    }
    shift[0] = shift[1] = shift[2] = 0.0;                               /* :97-102 */
    for (int i = 0; i < na; i++) {
        shift[0] += (avg[3 * i]     - (R[0] * x[3 * i] + R[1] * x[3 * i + 1] + R[2] * x[3 * i + 2]));
        shift[1] += (avg[3 * i + 1] - (R[3] * x[3 * i] + R[4] * x[3 * i + 1] + R[5] * x[3 * i + 2]));
        shift[2] += (avg[3 * i + 2] - (R[6] * x[3 * i] + R[7] * x[3 * i + 1] + R[8] * x[3 * i + 2]));
    }
    shift[0] /= na; shift[1] /= na; shift[2] /= na;

    for (int j = 0; j < nev; j++) {                                     /* :106-110 */
        double p = 0.0;
        for (int k = 0; k < n3; k++) p += ev[(size_t)j * n3 + k] * d[k];
        proj[j] = p;
    }
    for (int k = 0; k < n3; k++) {                                      /* :113-127 */
        double sh = avg[k], fk = 0.0;
        for (int j = 0; j < nev; j++) {
            const double e = ev[(size_t)j * n3 + k];
            sh += e * proj[j];
            fk -= (e * proj[j] * s->scaled / lam[j]);
        }
        d[k] = sh; f[k] = fk;
    }
    for (int i = 0; i < na; i++) {                                      /* :130-143 */
        d[3 * i] -= shift[0]; d[3 * i + 1] -= shift[1]; d[3 * i + 2] -= shift[2];
        x[3 * i]     = R[0] * d[3 * i] + R[3] * d[3 * i + 1] + R[6] * d[3 * i + 2];
        x[3 * i + 1] = R[1] * d[3 * i] + R[4] * d[3 * i + 1] + R[7] * d[3 * i + 2];
        x[3 * i + 2] = R[2] * d[3 * i] + R[5] * d[3 * i + 1] + R[8] * d[3 * i + 2];
        frot[3 * i]     = R[0] * f[3 * i] + R[3] * f[3 * i + 1] + R[6] * f[3 * i + 2];
        frot[3 * i + 1] = R[1] * f[3 * i] + R[4] * f[3 * i + 1] + R[7] * f[3 * i + 2];
        frot[3 * i + 2] = R[2] * f[3 * i] + R[5] * f[3 * i + 1] + R[8] * f[3 * i + 2];
    }
    memcpy(f, frot, sizeof(double) * n3);                               /* :146-150 */

    double e = 0.0;                                                     /* :153-157 */
    for (int j = 0; j < nev; j++) e += (proj[j] * proj[j] / lam[j]);
    e *= 0.5 * s->scaled;
    s->e_ed[t] = e;
    free(work);
}

/* ------------------------------------------------------------------------------------
 * glibc TYPE_3 rand()/srand()  (third-party; call sites edmd.cpp:179,206,207)
 * ---------------------------------------------------------------------------------- */
typedef struct { uint32_t r[31]; int f, b; } glibc_rand_t;

static uint32_t glibc_next(glibc_rand_t* g) {
    g->r[g->f] += g->r[g->b];
    const uint32_t out = (g->r[g->f] >> 1) & 0x7fffffffu;
    if (++g->f == 31) g->f = 0;
    if (++g->b == 31) g->b = 0;
    return out;
}
static void glibc_seed(glibc_rand_t* g, uint32_t seed) {
    if (seed == 0) seed = 1;
    g->r[0] = seed;
    int32_t w = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        const int32_t hi = w / 127773, lo = w % 127773;
        w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        g->r[i] = (uint32_t)w;
    }
    g->f = 3; g->b = 0;
    for (int i = 0; i < 310; i++) (void)glibc_next(g);
}
/* exported for tests/test_oracle.py (compares with libc) */
void orc_glibc_rand_stream(unsigned seed, int count, int* out) {
    glibc_rand_t g; glibc_seed(&g, seed);
    for (int i = 0; i < count; i++) out[i] = (int)glibc_next(&g);
}

/* EDMD::calculate_Random_Terms, edmd.cpp:170-228 */
static unsigned g_rng_seed = 13579u; /* function-static RNG_Seed, edmd.cpp:173 */
void orc_random_terms(orc_sys* s, int t, int rank) {
    const int n3 = 3 * s->na[t];
    const double* m = s->mass + s->off3[t];
    double* out = s->rnd + s->off3[t];
    glibc_rand_t g;
    glibc_seed(&g, (unsigned)((g_rng_seed++) + rank * rank * rank + g_time0));  /* :179 */
    if (g_rng_seed > 50000000u) g_rng_seed = 13579u;                               /* :180 */
    g_rng_calls++;
    for (int k = 0; k < n3; k++) {
        const double noise = sqrt(2.0 * s->gamma * s->scaled * m[k] / s->dt);      /* :184 */
        double u, v;
        for (;;) {                                                                 /* :204-217 */
            u = (double)(glibc_next(&g) / (double)2147483647);
            v = (double)(glibc_next(&g) / (double)2147483647);
            v = 1.7156 * (v - 0.5);
            const double x = u - 0.449871;
            const double y = fabs(v) - (-0.386595);
            const double q = x * x + y * (0.19600 * y - 0.25472 * x);
            if (q < 0.27597) break;
            if (q > 0.27846) continue;
            if (v * v < -4.0 * log(u) * (u * u)) break;
        }
        out[k] = (v / u) * noise;                                                  /* :220-223 */
    }
}

void orc_set_time(long t0) { g_time0 = t0; }
long orc_rng_calls(void) { return g_rng_calls; }
int orc_set_rng_calls(long calls) {
    /* replay the counter arithmetic of edmd.cpp:179-180 */
    g_rng_calls = calls;
    unsigned sd = 13579u;
    for (long c = 0; c < calls; c++) { sd++; if (sd > 50000000u) sd = 13579u; }
    g_rng_seed = sd;
    return 0;
}

/* ------------------------------------------------------------------------------------
 * EDMD::calculate_NB_Forces, edmd.cpp:232-285.  f1/f2 have 3N+2 entries each.
 * ---------------------------------------------------------------------------------- */
static void nb_pair(const orc_sys* s, int t1, int t2, double* f1, double* f2) {
    const int n1 = s->na[t1], n2 = s->na[t2];
    const double* x1 = s->crd + s->off3[t1];
    const double* x2 = s->crd + s->off3[t2];
    const double* q1 = s->abq + s->off3[t1];
    const double* q2 = s->abq + s->off3[t2];
    const double krep = s->krep, qfac = s->qfac;
    for (int k = 0; k < 3 * n1 + 2; k++) f1[k] = 0.0;
    for (int k = 0; k < 3 * n2 + 2; k++) f2[k] = 0.0;
    for (int i = 0; i < n1; i++) {
        for (int j = 0; j < n2; j++) {
            const double dx = x1[3 * i] - x2[3 * j];
            const double dy = x1[3 * i + 1] - x2[3 * j + 1];
            const double dz = x1[3 * i + 2] - x2[3 * j + 2];
            double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < 1e-9) r2 = 1e-9;                                   /* max(), :256 */
            if (r2 < (s->atom_cutoff * s->atom_cutoff)) {
                const double a = (2.0 - r2) > 0.0 ? (2.0 - r2) : 0.0;   /* max(0, 2-r2), :260 */
                const double qq = q1[3 * i + 2] * q2[3 * j + 2];
                f1[3 * n1]     += 0.25 * krep * a * a;
                f1[3 * n1 + 1] += 0.5 * qfac * qq * r2;
                const double pf = -2.0 * krep * a - 2.0 * qfac * qq / (r2 * r2);
                f1[3 * i] -= dx * pf; f1[3 * i + 1] -= dy * pf; f1[3 * i + 2] -= dz * pf;
                f2[3 * j] += dx * pf; f2[3 * j + 1] += dy * pf; f2[3 * j + 2] += dz * pf;
            }
        }
    }
    f2[3 * n2] = f1[3 * n1];                                            /* :282-283 */
    f2[3 * n2 + 1] = f1[3 * n1 + 1];
}

void orc_nb_forces(orc_sys* s, int t1, int t2) {
    const int stride = 3 * s->max_atoms + 2;
    double* f1 = (double*)malloc(sizeof(double) * 2 * stride);
    double* f2 = f1 + stride;
    nb_pair(s, t1, t2, f1, f2);
    memcpy(s->fnb + s->off3[t1], f1, 8 * 3 * (size_t)s->na[t1]);
    s->e_nb[2 * t1] = f1[3 * s->na[t1]]; s->e_nb[2 * t1 + 1] = f1[3 * s->na[t1] + 1];
    memcpy(s->fnb + s->off3[t2], f2, 8 * 3 * (size_t)s->na[t2]);
    s->e_nb[2 * t2] = f2[3 * s->na[t2]]; s->e_nb[2 * t2 + 1] = f2[3 * s->na[t2] + 1];
    free(f1);
}

/* EDMD::update_Velocities, edmd.cpp:289-317 */
void orc_update_velocities(orc_sys* s, int t) {
    const int na = s->na[t], n3 = 3 * na;
    double* v = s->vel + s->off3[t];
    const double* fe = s->fed + s->off3[t];
    const double* rn = s->rnd + s->off3[t];
    const double* fn = s->fnb + s->off3[t];
    const double* m = s->mass + s->off3[t];
    const double gamfac = 1.0 / (1.0 + s->gamma * s->dt);
    double ke = 0.0;
    for (int k = 0; k < n3; k++) {
        v[k] = (v[k] + fe[k] * s->dt + (rn[k] + fn[k]) * s->dt / m[k]) * gamfac;
        ke += 0.5 * m[k] * v[k] * v[k];
    }
    const double target = 0.5 * s->scaled * 3 * na;
    const double tscal = sqrt(1.0 + (s->dt / s->tautp) * ((target / ke) - 1.0));
    double T = ke * 2 / (KB_INTERNAL * 3 * na);
    T *= tscal * tscal;
    s->temp[t] = T;
    for (int k = 0; k < n3; k++) v[k] *= tscal;
}

/* EDMD::update_Coordinates, edmd.cpp:322-327 */
void orc_update_coordinates(orc_sys* s, int t) {
    const int n3 = 3 * s->na[t];
    double* x = s->crd + s->off3[t];
    const double* v = s->vel + s->off3[t];
    for (int k = 0; k < n3; k++) x[k] += v[k] * s->dt;
}

/* ------------------------------------------------------------------------------------
 * Pair list: Master::cal_Centre_of_Mass + generate_Pair_Lists, master.cpp:157-210
 * ---------------------------------------------------------------------------------- */
static void push_pair(orc_sys* s, int a, int b) {
    if (s->np == s->pairs_cap) {
        s->pairs_cap = s->pairs_cap ? 2 * s->pairs_cap : 1024;
        s->pairs = (int*)realloc(s->pairs, sizeof(int) * 2 * (size_t)s->pairs_cap);
    }
    s->pairs[2 * s->np] = a; s->pairs[2 * s->np + 1] = b; s->np++;
}

long long orc_build_pairs(orc_sys* s) {
    const int n = s->n;
    double* c = (double*)malloc(sizeof(double) * 3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        const double* x = s->crd + s->off3[i];
        double cx = 0.0, cy = 0.0, cz = 0.0;
        for (int j = 0; j < 3 * s->na[i]; j += 3) { cx += x[j]; cy += x[j + 1]; cz += x[j + 2]; }
        c[3 * i] = cx / s->na[i]; c[3 * i + 1] = cy / s->na[i]; c[3 * i + 2] = cz / s->na[i];
    }
    s->np = 0;
    const double cut2 = s->mole_cutoff * s->mole_cutoff;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            const double r = (c[3*i] - c[3*j]) * (c[3*i] - c[3*j])
                           + (c[3*i+1] - c[3*j+1]) * (c[3*i+1] - c[3*j+1])
                           + (c[3*i+2] - c[3*j+2]) * (c[3*i+2] - c[3*j+2]);
            const int sep = j - i;
            if ((r < cut2) && (sep > s->mole_least) && (sep < (n - s->mole_least))) {
                if (sep % 2 == 1) push_pair(s, j, i); else push_pair(s, i, j);   /* :197-201 */
            }
        }
    free(c);
    return s->np;
}
long long orc_num_pairs(orc_sys* s) { return s->np; }
void orc_get_pairs(orc_sys* s, int* out) { if (s->np) memcpy(out, s->pairs, sizeof(int) * 2 * (size_t)s->np); }
void orc_set_pairs(orc_sys* s, long long np, const int* pairs) {
    s->np = 0;
    for (long long p = 0; p < np; p++) push_pair(s, pairs[2 * p], pairs[2 * p + 1]);
}

/* ------------------------------------------------------------------------------------
 * Force evaluation protocol: worker.cpp:147-199 (W = 1) + master.cpp:297-322
 * ---------------------------------------------------------------------------------- */
static void split(long long total, int workers, int w, long long* lo, long long* hi) {
    /* Master::generate_Indexes, master.cpp:214-242 */
    const long long base = total / workers, rem = total - base * workers;
    *lo = w * base + (w < rem ? w : rem);
    *hi = *lo + base + (w < rem ? 1 : 0);
}

static void nb_range(const orc_sys* s, long long p0, long long p1, double* acc, int stride) {
    double* f1 = (double*)malloc(sizeof(double) * 2 * stride);
    double* f2 = f1 + stride;
    for (long long p = p0; p < p1; p++) {
        const int t1 = s->pairs[2 * p], t2 = s->pairs[2 * p + 1];
        nb_pair(s, t1, t2, f1, f2);
        double* r1 = acc + (size_t)t1 * stride;
        double* r2 = acc + (size_t)t2 * stride;
        for (int k = 0; k < 3 * s->na[t1] + 2; k++) r1[k] += f1[k];   /* worker.cpp:173-178 */
        for (int k = 0; k < 3 * s->na[t2] + 2; k++) r2[k] += f2[k];
    }
    free(f1);
}

typedef struct { orc_sys* s; long long lo, hi; double* acc; int stride; int what; } job_t;
static void* job_main(void* arg) {
    job_t* j = (job_t*)arg;
    if (j->what == 0) for (long long t = j->lo; t < j->hi; t++) orc_ed_forces(j->s, (int)t);
    else nb_range(j->s, j->lo, j->hi, j->acc, j->stride);
    return NULL;
}

static void clip_and_store(orc_sys* s, const double* acc, int stride) { /* master.cpp:297-322 */
    for (int t = 0; t < s->n; t++) {
        const int n3 = 3 * s->na[t];
        const double* r = acc + (size_t)t * stride;
        double* f = s->fnb + s->off3[t];
        for (int k = 0; k < n3; k++) {
            double v = r[k];
            if (v < -1.0) v = -1.0; else if (v > 1.0) v = 1.0;
            f[k] = v;
        }
        s->e_nb[2 * t] = r[n3]; s->e_nb[2 * t + 1] = r[n3 + 1];
    }
}

void orc_forces(orc_sys* s, int random_on, int nthreads) {
    const int n = s->n, stride = 3 * s->max_atoms + 2;
    if (nthreads < 1) nthreads = 1;
    double* acc = (double*)calloc((size_t)n * stride, 8);
    if (nthreads == 1) {
        for (int t = 0; t < n; t++) {
            orc_ed_forces(s, t);
            if (random_on) orc_random_terms(s, t, 1);
            else memset(s->rnd + s->off3[t], 0, 8 * 3 * (size_t)s->na[t]);
        }
        nb_range(s, 0, s->np, acc, stride);
    } else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
        job_t* jb = (job_t*)malloc(sizeof(job_t) * nthreads);
        for (int w = 0; w < nthreads; w++) {
            jb[w].s = s; jb[w].what = 0; split(n, nthreads, w, &jb[w].lo, &jb[w].hi);
            pthread_create(&th[w], NULL, job_main, &jb[w]);
        }
        for (int w = 0; w < nthreads; w++) pthread_join(th[w], NULL);
        for (int t = 0; t < n; t++) {
            if (random_on) orc_random_terms(s, t, 1);
            else memset(s->rnd + s->off3[t], 0, 8 * 3 * (size_t)s->na[t]);
        }
        double* part = (double*)calloc((size_t)nthreads * n * stride, 8);
        for (int w = 0; w < nthreads; w++) {
            jb[w].what = 1; jb[w].acc = part + (size_t)w * n * stride; jb[w].stride = stride;
            split(s->np, nthreads, w, &jb[w].lo, &jb[w].hi);
            pthread_create(&th[w], NULL, job_main, &jb[w]);
        }
        for (int w = 0; w < nthreads; w++) pthread_join(th[w], NULL);
        for (int w = 0; w < nthreads; w++)
            for (size_t k = 0; k < (size_t)n * stride; k++) acc[k] += part[(size_t)w * n * stride + k];
        free(part); free(jb); free(th);
    }
    clip_and_store(s, acc, stride);
    free(acc);
}

void orc_step(orc_sys* s, int nsteps, int random_on, int nthreads) { /* simulation.cpp:42-48 */
    for (int it = 0; it < nsteps; it++) {
        orc_forces(s, random_on, nthreads);
        for (int t = 0; t < s->n; t++) orc_update_velocities(s, t);
        for (int t = 0; t < s->n; t++) orc_update_coordinates(s, t);
    }
}

double orc_time_nb(orc_sys* s, long long p0, long long p1, int nthreads) {
    const int n = s->n, stride = 3 * s->max_atoms + 2;
    if (nthreads < 1) nthreads = 1;
    if (p1 > s->np) p1 = s->np;
    double* part = (double*)calloc((size_t)nthreads * n * stride, 8);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
    job_t* jb = (job_t*)malloc(sizeof(job_t) * nthreads);
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int w = 0; w < nthreads; w++) {
        jb[w].s = s; jb[w].what = 1; jb[w].acc = part + (size_t)w * n * stride; jb[w].stride = stride;
        split(p1 - p0, nthreads, w, &jb[w].lo, &jb[w].hi);
        jb[w].lo += p0; jb[w].hi += p0;
        pthread_create(&th[w], NULL, job_main, &jb[w]);
    }
    for (int w = 0; w < nthreads; w++) pthread_join(th[w], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    free(part); free(jb); free(th);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

/* Master::merge_Vels_n_Crds, master.cpp:347-384 (displacements: io.cpp:236-247) */
void orc_merge(orc_sys* s, const int* bp_atoms) {
    const int n = s->n, nbp = n + 3;
    long long* dsp = (long long*)calloc(nbp + 1, sizeof(long long));
    for (int i = 0; i < nbp; i++) dsp[i + 1] = dsp[i] + 3LL * bp_atoms[i];
    const long long tot = dsp[nbp];
    double* V = (double*)calloc((size_t)tot, 8);
    double* X = (double*)calloc((size_t)tot, 8);
    for (int t = 0; t < n; t++) {
        const double* v = s->vel + s->off3[t];
        const double* x = s->crd + s->off3[t];
        for (int k = 0; k < 3 * s->na[t]; k++) { V[dsp[t] + k] += v[k]; X[dsp[t] + k] += x[k]; }
    }
    for (long long hi = dsp[nbp - 3], lo = 0; hi < dsp[nbp]; hi++, lo++) {  /* ring closure :365-369 */
        V[hi] += V[lo]; V[lo] = V[hi];
        X[hi] += X[lo]; X[lo] = X[hi];
    }
    for (long long k = 0; k < tot; k++) { V[k] *= 0.25; X[k] *= 0.25; }
    for (int t = 0; t < n; t++) {
        double* v = s->vel + s->off3[t];
        double* x = s->crd + s->off3[t];
        for (int k = 0; k < 3 * s->na[t]; k++) { v[k] = V[dsp[t] + k]; x[k] = X[dsp[t] + k]; }
    }
    free(V); free(X); free(dsp);
}

/* Master::write_Info, master.cpp:388-401 */
void orc_energies(orc_sys* s, double e[4]) {
    e[0] = e[1] = e[2] = e[3] = 0.0;
    for (int t = 0; t < s->n; t++) {
        e[0] += s->e_ed[t]; e[1] += s->e_nb[2 * t]; e[2] += s->e_nb[2 * t + 1]; e[3] += s->temp[t];
    }
    e[3] /= s->n;
}
