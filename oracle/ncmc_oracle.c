/*
 * CPU ORACLE, C twin (test infrastructure and the bench's CPU baseline — NOT product code).
 *
 * Float64 + OpenMP restatement of the same path as oracle/ncmc_oracle.py, fast enough for the 22k-atom
 * T4 lysozyme workload.  It executes the REFERENCE's step program literally (blues/integrators.py:159-231):
 * `energy` is re-evaluated in full wherever the CustomIntegrator program reads it, i.e. >= 3 full force/energy
 * evaluations per NCMC step — this is what BLUES + OpenMM do, and what the "host CPU" baseline measures.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Arithmetic restated from the pinned third-party dependencies (openmmtools 0.15.0, OpenMM 7.x), which are
 * not part of the reference checkout; see the header of ncmc_oracle.py for the list and for the parity-pinning
 * statement (force/energy arithmetic: parity unpinned against OpenMM; validated against ncmc_oracle.py).
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/blues_b200.h"

#define KE_COUL 138.935456
#define KB_KJ (1.3806504e-23 * 6.02214179e23 / 1000.0)
#define PME_ORDER 5
#define TWO_OVER_SQRT_PI 1.1283791670955126

typedef struct { double re, im; } cplx;

typedef struct orc_handle {
    bl_topology t;              /* deep copy */
    int N, periodic, pme;
    double box[3];
    double* invm;
    /* exclusions CSR */
    int* ex_ptr; int* ex_idx;
    unsigned char* is_alch; double *aq, *asig, *aeps;   /* per-atom alchemical parameters (0 elsewhere) */
    /* verlet list */
    int* nl_ptr; int* nl_idx; long nl_cap; double* x_ref; double skin; int nl_valid;
    /* pme */
    int K[3]; double* Q; cplx* S; double* bmod[3];
    /* integrator */
    char split[32]; int nsplit;
    double kT, gamma, dt, tol;
    int nsteps, nprop, n_lambda_steps, nV, nR, nO;
    double pl_min, pl_max;
    double *lam_s_tab, *lam_e_tab;
    uint64_t seed; uint32_t replica;
    /* globals */
    double protocol_work, perturbed_pe, unperturbed_pe, lambda, Eold, Enew, heat;
    int step, lambda_step, first_step, prop;
    uint32_t noise_counter, vel_counter;
    double lam_s, lam_e;
    long n_evals;
    double* F;                  /* scratch forces */
    int nthreads;
    double* Fthr;               /* per-thread force buffers */
    int cache_valid; double cache_E; double* cache_F;   /* evaluation at the current (x, lambda) */
    /* "optimised CPU" mode (bench context row, not the reference's semantics): one full evaluation per coordinate set,
     * lambda changes re-evaluate only the pairs that involve alchemical atoms (the lambda-separable form the engine uses) */
    int fast, xcache_valid; double env_E; double* env_F; double* alch_F;
    int* apairs; long n_apairs, apairs_cap;
} orc_handle;

/* ------------------------------------------------------------------------------------------------ philox */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }
static void normal3(uint64_t seed, uint32_t stream, uint32_t rep, uint32_t counter, uint32_t idx, double n[3]) {
    uint32_t r[4];
    philox(idx, counter, rep, stream, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    double r0 = sqrt(-2.0 * log(u01(r[0]))), r1 = sqrt(-2.0 * log(u01(r[2])));
    n[0] = r0 * cos(2 * M_PI * u01(r[1]));
    n[1] = r0 * sin(2 * M_PI * u01(r[1]));
    n[2] = r1 * cos(2 * M_PI * u01(r[3]));
}

/* ------------------------------------------------------------------------------------------------ helpers */
static inline void minimg(const orc_handle* h, double d[3]) {
    if (h->periodic) for (int k = 0; k < 3; ++k) d[k] -= h->box[k] * rint(d[k] / h->box[k]);
}
static void* dup_mem(const void* p, size_t n) { void* q = malloc(n ? n : 1); if (n) memcpy(q, p, n); return q; }

static double softcore(double r, double sig, double eps, double lam, double alpha, double a, double b, double c, double* f_over_r) {
    double rs = r / sig, la = pow(lam, a), s = alpha * pow(1.0 - lam, b) + pow(rs, c);
    double x = pow(s, -6.0 / c);
    double dx = (-6.0 / c) * pow(s, -6.0 / c - 1.0) * c * pow(rs, c - 1.0) / sig;
    *f_over_r = -(la * 4.0 * eps * (2.0 * x - 1.0) * dx) / r;
    return la * 4.0 * eps * x * (x - 1.0);
}

/* ------------------------------------------------------------------------------------------------ FFT */
static void dft_line(cplx* a, int n, int stride, int sign, cplx* tmp) {
    /* mixed-radix decimation in time on a strided line, recursive, radices 2,3,5,7 (generic prime fallback) */
    if (n == 1) return;
    int p = 0;
    for (int q = 2; q <= n; ++q) if (n % q == 0) { p = q; break; }
    int m = n / p;
    /* split into p interleaved sub-sequences, transform each */
    for (int r = 0; r < p; ++r) dft_line(a + (size_t)r * stride, m, stride * p, sign, tmp);
    for (int k = 0; k < n; ++k) tmp[k] = a[(size_t)k * stride];
    for (int k = 0; k < m; ++k) {
        for (int q = 0; q < p; ++q) {
            /* output index k + q m */
            double sr = 0, si = 0;
            for (int r = 0; r < p; ++r) {
                double ang = sign * 2.0 * M_PI * (double)r * (double)(k + q * m) / (double)n;
                double c = cos(ang), s = sin(ang);
                cplx v = tmp[k * p + r];      /* element k of sub-sequence r sits at original index k p + r */
                sr += v.re * c - v.im * s;
                si += v.re * s + v.im * c;
            }
            a[(size_t)(k + q * m) * stride].re = sr;
            a[(size_t)(k + q * m) * stride].im = si;
        }
    }
}
static void fft3(cplx* g, const int K[3], int sign) {
    const int nx = K[0], ny = K[1], nz = K[2];
#pragma omp parallel
    {
        cplx* tmp = (cplx*)malloc(sizeof(cplx) * (size_t)(nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz)));
#pragma omp for collapse(2)
        for (int i = 0; i < nx; ++i) for (int j = 0; j < ny; ++j) dft_line(g + ((size_t)i * ny + j) * nz, nz, 1, sign, tmp);
#pragma omp for collapse(2)
        for (int i = 0; i < nx; ++i) for (int k = 0; k < nz; ++k) dft_line(g + (size_t)i * ny * nz + k, ny, nz, sign, tmp);
#pragma omp for collapse(2)
        for (int j = 0; j < ny; ++j) for (int k = 0; k < nz; ++k) dft_line(g + (size_t)j * nz + k, nx, ny * nz, sign, tmp);
        free(tmp);
    }
}

static double M_spline(int order, double u) {
    double out = 0, fact = 1;
    for (int k = 1; k < order; ++k) fact *= k;
    double binom = 1;
    for (int k = 0; k <= order; ++k) {
        double t = u - k;
        if (t > 0) out += ((k & 1) ? -1.0 : 1.0) * binom * pow(t, order - 1);
        binom = binom * (order - k) / (k + 1);
    }
    return out / fact;
}
static void bspline(double w, double* v, double* dv) {
    /* a[k] = M_n(w + k) built by the standard recursion; v[k] = M5(w + 4 - k) */
    double a[PME_ORDER] = {w, 1.0 - w, 0, 0, 0}, da[PME_ORDER];
    for (int n = 3; n <= PME_ORDER; ++n) {
        if (n == PME_ORDER) { da[0] = a[0]; for (int k = 1; k < PME_ORDER - 1; ++k) da[k] = a[k] - a[k - 1]; da[PME_ORDER - 1] = -a[PME_ORDER - 2]; }
        for (int k = n - 1; k >= 0; --k) {
            double u = w + k, lo = (k < n - 1) ? a[k] : 0.0, hi = (k > 0) ? a[k - 1] : 0.0;
            a[k] = (u * lo + (n - u) * hi) / (n - 1);
        }
    }
    for (int k = 0; k < PME_ORDER; ++k) { v[k] = a[PME_ORDER - 1 - k]; dv[k] = da[PME_ORDER - 1 - k]; }
}

static double pme_reciprocal(orc_handle* h, const double* x, double* F) {
    const int* K = h->K;
    const size_t G = (size_t)K[0] * K[1] * K[2];
    const double* q = h->t.charge;
    memset(h->Q, 0, sizeof(double) * G);
    const int N = h->N;
    /* spread (serial: deterministic) */
    for (int a = 0; a < N; ++a) {
        if (q[a] == 0) continue;
        int base[3]; double w[3][PME_ORDER], dw[3][PME_ORDER];
        for (int d = 0; d < 3; ++d) {
            double f = x[3 * a + d] / h->box[d]; f -= floor(f);
            double u = f * K[d]; int b = (int)floor(u);
            bspline(u - b, w[d], dw[d]);
            base[d] = b % K[d];
        }
        for (int i = 0; i < PME_ORDER; ++i) for (int j = 0; j < PME_ORDER; ++j) for (int k = 0; k < PME_ORDER; ++k) {
            int gx = (base[0] + i) % K[0], gy = (base[1] + j) % K[1], gz = (base[2] + k) % K[2];
            h->Q[((size_t)gx * K[1] + gy) * K[2] + gz] += q[a] * w[0][i] * w[1][j] * w[2][k];
        }
    }
    for (size_t i = 0; i < G; ++i) { h->S[i].re = h->Q[i]; h->S[i].im = 0; }
    fft3(h->S, K, -1);
    const double V = h->box[0] * h->box[1] * h->box[2], alpha = h->t.ewald_alpha;
    double E = 0;
#pragma omp parallel for reduction(+ : E)
    for (int i = 0; i < K[0]; ++i) for (int j = 0; j < K[1]; ++j) for (int k = 0; k < K[2]; ++k) {
        size_t idx = ((size_t)i * K[1] + j) * K[2] + k;
        if (idx == 0) { h->S[0].re = h->S[0].im = 0; continue; }
        int mi = i <= K[0] / 2 ? i : i - K[0], mj = j <= K[1] / 2 ? j : j - K[1], mk = k <= K[2] / 2 ? k : k - K[2];
        double fx = mi / h->box[0], fy = mj / h->box[1], fz = mk / h->box[2];
        double m2 = fx * fx + fy * fy + fz * fz;
        double Gf = exp(-M_PI * M_PI * m2 / (alpha * alpha)) / (m2 * h->bmod[0][i] * h->bmod[1][j] * h->bmod[2][k] * M_PI * V);
        E += 0.5 * KE_COUL * Gf * (h->S[idx].re * h->S[idx].re + h->S[idx].im * h->S[idx].im);
        h->S[idx].re *= Gf * KE_COUL; h->S[idx].im *= Gf * KE_COUL;
    }
    fft3(h->S, K, +1);     /* unnormalised inverse: potential on the grid */
#pragma omp parallel for
    for (int a = 0; a < N; ++a) {
        if (q[a] == 0) continue;
        int base[3]; double w[3][PME_ORDER], dw[3][PME_ORDER];
        for (int d = 0; d < 3; ++d) {
            double f = x[3 * a + d] / h->box[d]; f -= floor(f);
            double u = f * K[d]; int b = (int)floor(u);
            bspline(u - b, w[d], dw[d]);
            base[d] = b % K[d];
        }
        double f[3] = {0, 0, 0};
        for (int i = 0; i < PME_ORDER; ++i) for (int j = 0; j < PME_ORDER; ++j) for (int k = 0; k < PME_ORDER; ++k) {
            int gx = (base[0] + i) % K[0], gy = (base[1] + j) % K[1], gz = (base[2] + k) % K[2];
            double phi = h->S[((size_t)gx * K[1] + gy) * K[2] + gz].re;
            f[0] += phi * dw[0][i] * w[1][j] * w[2][k];
            f[1] += phi * w[0][i] * dw[1][j] * w[2][k];
            f[2] += phi * w[0][i] * w[1][j] * dw[2][k];
        }
        for (int d = 0; d < 3; ++d) F[3 * a + d] -= q[a] * f[d] * K[d] / h->box[d];
    }
    return E;
}

/* ------------------------------------------------------------------------------------------------ neighbour list */
static int excluded(const orc_handle* h, int i, int j) {
    for (int k = h->ex_ptr[i]; k < h->ex_ptr[i + 1]; ++k) if (h->ex_idx[k] == j) return 1;
    return 0;
}
static void build_list(orc_handle* h, const double* x) {
    const int N = h->N;
    const double rl = h->periodic ? h->t.cutoff + h->skin : 1e30, rl2 = rl * rl;
    int nc[3] = {1, 1, 1};
    if (h->periodic) for (int d = 0; d < 3; ++d) { nc[d] = (int)floor(h->box[d] / rl); if (nc[d] < 1) nc[d] = 1; }
    const int ncell = nc[0] * nc[1] * nc[2];
    int* head = (int*)malloc(sizeof(int) * (ncell + 1));
    int* cell = (int*)malloc(sizeof(int) * N);
    int* order = (int*)malloc(sizeof(int) * N);
    memset(head, 0, sizeof(int) * (ncell + 1));
    for (int a = 0; a < N; ++a) {
        int c[3] = {0, 0, 0};
        if (h->periodic) for (int d = 0; d < 3; ++d) {
            double f = x[3 * a + d] / h->box[d]; f -= floor(f);
            c[d] = (int)(f * nc[d]); if (c[d] >= nc[d]) c[d] = nc[d] - 1;
        }
        cell[a] = (c[0] * nc[1] + c[1]) * nc[2] + c[2];
        head[cell[a] + 1]++;
    }
    for (int c = 0; c < ncell; ++c) head[c + 1] += head[c];
    int* fill = (int*)dup_mem(head, sizeof(int) * (ncell + 1));
    for (int a = 0; a < N; ++a) order[fill[cell[a]]++] = a;
    free(fill);
    /* count then fill: neighbours j > i only */
    int* cnt = (int*)calloc(N + 1, sizeof(int));
    for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < N; ++i) {
            int ci = cell[i];
            int cz = ci % nc[2], cy = (ci / nc[2]) % nc[1], cx = ci / (nc[1] * nc[2]);
            int n = 0;
            int* dst = pass ? h->nl_idx + h->nl_ptr[i] : NULL;
            int seen[27], nseen = 0;
            for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz) {
                int ax = cx + dx, ay = cy + dy, az = cz + dz;
                if (h->periodic) { ax = (ax + nc[0]) % nc[0]; ay = (ay + nc[1]) % nc[1]; az = (az + nc[2]) % nc[2]; }
                else if (ax || ay || az) continue;
                int cj = (ax * nc[1] + ay) * nc[2] + az, dup = 0;
                for (int s = 0; s < nseen; ++s) if (seen[s] == cj) dup = 1;
                if (dup) continue;
                seen[nseen++] = cj;
                for (int k = head[cj]; k < head[cj + 1]; ++k) {
                    int j = order[k];
                    if (j <= i) continue;
                    double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
                    minimg(h, d);
                    if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] >= rl2) continue;
                    if (excluded(h, i, j)) continue;
                    if (pass) dst[n] = j;
                    n++;
                }
            }
            if (!pass) cnt[i + 1] = n;
        }
        if (!pass) {
            for (int i = 0; i < N; ++i) cnt[i + 1] += cnt[i];
            memcpy(h->nl_ptr, cnt, sizeof(int) * (N + 1));
            if (cnt[N] > h->nl_cap) { h->nl_cap = (long)(cnt[N] * 1.2) + 1024; h->nl_idx = (int*)realloc(h->nl_idx, sizeof(int) * h->nl_cap); }
        }
    }
    free(cnt); free(head); free(cell); free(order);
    memcpy(h->x_ref, x, sizeof(double) * 3 * N);
    h->nl_valid = 1;
    if (h->fast) {
        h->n_apairs = 0;
        for (int i = 0; i < N; ++i)
            for (int k = h->nl_ptr[i]; k < h->nl_ptr[i + 1]; ++k) {
                int j = h->nl_idx[k];
                if (!h->is_alch[i] && !h->is_alch[j]) continue;
                if (h->n_apairs + 1 > h->apairs_cap) { h->apairs_cap = h->apairs_cap * 2 + 1024; h->apairs = (int*)realloc(h->apairs, sizeof(int) * 2 * h->apairs_cap); }
                h->apairs[2 * h->n_apairs] = i; h->apairs[2 * h->n_apairs + 1] = j; h->n_apairs++;
            }
    }
}
static void ensure_list(orc_handle* h, const double* x) {
    if (h->nl_valid && h->periodic) {
        double lim = 0.25 * h->skin * h->skin; int moved = 0;
#pragma omp parallel for reduction(| : moved)
        for (int a = 0; a < h->N; ++a) {
            double d0 = x[3 * a] - h->x_ref[3 * a], d1 = x[3 * a + 1] - h->x_ref[3 * a + 1], d2 = x[3 * a + 2] - h->x_ref[3 * a + 2];
            if (d0 * d0 + d1 * d1 + d2 * d2 > lim) moved = 1;
        }
        if (!moved) return;
    } else if (h->nl_valid) {
        return;     /* non-periodic all-pairs list never changes */
    }
    build_list(h, x);
}

/* ------------------------------------------------------------------------------------------------ energy + forces */
double orc_energy_forces(orc_handle* h, const double* x, double lam_s, double lam_e, double* Fout, double* terms /*[12] or NULL*/) {
    const bl_topology* t = &h->t;
    const int N = h->N;
    double* F = Fout ? Fout : h->F;
    memset(F, 0, sizeof(double) * 3 * N);
    double e_bond = 0, e_angle = 0, e_tors = 0, e_restr = 0, e_pair = 0, e_exc = 0, e_pme = 0, e_ast = 0, e_ael = 0, e_aex = 0;
    h->n_evals++;
    /* bonded (serial: small) */
    for (int k = 0; k < t->n_bonds; ++k) {
        int i = t->bonds[2 * k], j = t->bonds[2 * k + 1];
        double d[3] = {x[3 * j] - x[3 * i], x[3 * j + 1] - x[3 * i + 1], x[3 * j + 2] - x[3 * i + 2]};
        minimg(h, d);
        double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), dr = r - t->bond_r0[k];
        e_bond += 0.5 * t->bond_k[k] * dr * dr;
        double c = t->bond_k[k] * dr / r;
        for (int q = 0; q < 3; ++q) { F[3 * i + q] += c * d[q]; F[3 * j + q] -= c * d[q]; }
    }
    for (int k = 0; k < t->n_angles; ++k) {
        int a = t->angles[3 * k], b = t->angles[3 * k + 1], c = t->angles[3 * k + 2];
        double v1[3], v2[3];
        for (int q = 0; q < 3; ++q) { v1[q] = x[3 * a + q] - x[3 * b + q]; v2[q] = x[3 * c + q] - x[3 * b + q]; }
        minimg(h, v1); minimg(h, v2);
        double r1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]), r2 = sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
        double cs = (v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) / (r1 * r2);
        cs = cs > 1 ? 1 : (cs < -1 ? -1 : cs);
        double th = acos(cs), dth = th - t->angle_t0[k];
        e_angle += 0.5 * t->angle_k[k] * dth * dth;
        double sn = sqrt(fmax(1 - cs * cs, 1e-30)), cf = t->angle_k[k] * dth / sn;
        for (int q = 0; q < 3; ++q) {
            double f1 = cf * (v2[q] / (r1 * r2) - cs * v1[q] / (r1 * r1)), f3 = cf * (v1[q] / (r1 * r2) - cs * v2[q] / (r2 * r2));
            F[3 * a + q] += f1; F[3 * c + q] += f3; F[3 * b + q] -= f1 + f3;
        }
    }
    for (int k = 0; k < t->n_torsions; ++k) {
        const int* ix = t->torsions + 4 * k;
        double b1[3], b2[3], b3[3];
        for (int q = 0; q < 3; ++q) { b1[q] = x[3 * ix[1] + q] - x[3 * ix[0] + q]; b2[q] = x[3 * ix[2] + q] - x[3 * ix[1] + q]; b3[q] = x[3 * ix[3] + q] - x[3 * ix[2] + q]; }
        minimg(h, b1); minimg(h, b2); minimg(h, b3);
        double n1[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
        double n2[3] = {b2[1] * b3[2] - b2[2] * b3[1], b2[2] * b3[0] - b2[0] * b3[2], b2[0] * b3[1] - b2[1] * b3[0]};
        double b2n = sqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
        double u[3] = {b2[0] / b2n, b2[1] / b2n, b2[2] / b2n};
        double m1[3] = {n1[1] * u[2] - n1[2] * u[1], n1[2] * u[0] - n1[0] * u[2], n1[0] * u[1] - n1[1] * u[0]};
        double xx = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2], yy = m1[0] * n2[0] + m1[1] * n2[1] + m1[2] * n2[2];
        double phi = atan2(yy, xx), arg = t->torsion_n[k] * phi - t->torsion_phase[k];
        e_tors += t->torsion_k[k] * (1 + cos(arg));
        double dE = -t->torsion_k[k] * t->torsion_n[k] * sin(arg);
        double n1s = n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2], n2s = n2[0] * n2[0] + n2[1] * n2[1] + n2[2] * n2[2];
        double s12 = (b1[0] * b2[0] + b1[1] * b2[1] + b1[2] * b2[2]) / (b2n * b2n), s32 = (b3[0] * b2[0] + b3[1] * b2[1] + b3[2] * b2[2]) / (b2n * b2n);
        for (int q = 0; q < 3; ++q) {
            double g0 = -b2n / n1s * n1[q], g3 = b2n / n2s * n2[q];
            double g1 = (-1 - s12) * g0 + s32 * g3, g2 = (-1 - s32) * g3 + s12 * g0;
            F[3 * ix[0] + q] += dE * g0; F[3 * ix[1] + q] += dE * g1; F[3 * ix[2] + q] += dE * g2; F[3 * ix[3] + q] += dE * g3;
        }
    }
    for (int k = 0; k < t->n_restraints; ++k) {
        int a = t->restraint_atoms[k];
        double d[3] = {x[3 * a] - t->restraint_x0[3 * k], x[3 * a + 1] - t->restraint_x0[3 * k + 1], x[3 * a + 2] - t->restraint_x0[3 * k + 2]};
        minimg(h, d);
        e_restr += t->restraint_k[k] * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int q = 0; q < 3; ++q) F[3 * a + q] -= 2 * t->restraint_k[k] * d[q];
    }
    /* exceptions + Ewald exclusion corrections */
    const double alpha = t->ewald_alpha;
    for (int k = 0; k < t->n_excl; ++k) {
        int i = t->excl_pairs[2 * k], j = t->excl_pairs[2 * k + 1];
        double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        minimg(h, d);
        double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], r = sqrt(r2), fr = 0;
        if (t->excl_eps[k] != 0) { double s2 = t->excl_sigma[k] * t->excl_sigma[k] / r2, s6 = s2 * s2 * s2; e_exc += 4 * t->excl_eps[k] * s6 * (s6 - 1); fr += 4 * t->excl_eps[k] * (12 * s6 * s6 - 6 * s6) / r2; }
        if (t->excl_qq[k] != 0) { double kq = KE_COUL * t->excl_qq[k]; e_exc += kq / r; fr += kq / (r * r2); }
        if (h->pme) {
            double kqq = KE_COUL * t->charge[i] * t->charge[j];
            if (kqq != 0) { double ar = alpha * r, er = erf(ar); e_exc -= kqq * er / r; fr += kqq * (TWO_OVER_SQRT_PI * alpha * exp(-ar * ar) / r - er / r2) / r; }
        }
        for (int q = 0; q < 3; ++q) { F[3 * i + q] += fr * d[q]; F[3 * j + q] -= fr * d[q]; }
    }
    for (int k = 0; k < t->n_alch_exc; ++k) {
        int i = t->alch_exc_pairs[2 * k], j = t->alch_exc_pairs[2 * k + 1];
        int both = h->is_alch[i] && h->is_alch[j];
        double ls = (both && !t->annihilate_sterics) ? 1.0 : lam_s, le = (both && !t->annihilate_electrostatics) ? 1.0 : lam_e;
        double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        minimg(h, d);
        double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], r = sqrt(r2), fr = 0;
        if (t->alch_exc_eps[k] != 0) e_aex += softcore(r, t->alch_exc_sigma[k], t->alch_exc_eps[k], ls, t->softcore_alpha, t->softcore_a, t->softcore_b, t->softcore_c, &fr);
        double kq = KE_COUL * t->alch_exc_qq[k] * le;
        e_aex += kq / r; fr += kq / (r * r2);
        for (int q = 0; q < 3; ++q) { F[3 * i + q] += fr * d[q]; F[3 * j + q] -= fr * d[q]; }
    }
    /* pair loop over the Verlet list */
    ensure_list(h, x);
    const double rc2 = h->periodic ? t->cutoff * t->cutoff : 1e300;
    const double krf = (t->nb_method == 2) ? (1.0 / (t->cutoff * t->cutoff * t->cutoff)) * (78.3 - 1) / (2 * 78.3 + 1) : 0;
    const double crf = (t->nb_method == 2) ? (1.0 / t->cutoff) * 3 * 78.3 / (2 * 78.3 + 1) : 0;
    memset(h->Fthr, 0, sizeof(double) * 3 * N * h->nthreads);
#pragma omp parallel reduction(+ : e_pair, e_ast, e_ael)
    {
#ifdef _OPENMP
        double* Ft = h->Fthr + (size_t)omp_get_thread_num() * 3 * N;
#else
        double* Ft = h->Fthr;
#endif
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < N; ++i) {
            const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
            const double qi = t->charge[i], si = t->sigma[i], ei = t->epsilon[i];
            double fi[3] = {0, 0, 0};
            for (int k = h->nl_ptr[i]; k < h->nl_ptr[i + 1]; ++k) {
                const int j = h->nl_idx[k];
                double d[3] = {xi - x[3 * j], yi - x[3 * j + 1], zi - x[3 * j + 2]};
                minimg(h, d);
                const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (r2 >= rc2) continue;
                const double r = sqrt(r2);
                double fr = 0;
                const double eps = sqrt(ei * t->epsilon[j]);
                if (eps > 0) { double sg = 0.5 * (si + t->sigma[j]), s2 = sg * sg / r2, s6 = s2 * s2 * s2; e_pair += 4 * eps * s6 * (s6 - 1); fr += 4 * eps * (12 * s6 * s6 - 6 * s6) / r2; }
                const double kqq = KE_COUL * qi * t->charge[j];
                if (kqq != 0) {
                    if (h->pme) { double ar = alpha * r, erc = erfc(ar); e_pair += kqq * erc / r; fr += kqq * (erc / r + TWO_OVER_SQRT_PI * alpha * exp(-ar * ar)) / r2; }
                    else if (t->nb_method == 2) { e_pair += kqq * (1 / r + krf * r2 - crf); fr += kqq * (1 / (r * r2) - 2 * krf); }
                    else { e_pair += kqq / r; fr += kqq / (r * r2); }
                }
                if (h->is_alch[i] || h->is_alch[j]) {
                    int both = h->is_alch[i] && h->is_alch[j];
                    double sa = 0.5 * ((h->is_alch[i] ? h->asig[i] : si) + (h->is_alch[j] ? h->asig[j] : t->sigma[j]));
                    double ea = sqrt((h->is_alch[i] ? h->aeps[i] : ei) * (h->is_alch[j] ? h->aeps[j] : t->epsilon[j]));
                    double qa = (h->is_alch[i] ? h->aq[i] : qi) * (h->is_alch[j] ? h->aq[j] : t->charge[j]);
                    double ls = (both && !t->annihilate_sterics) ? 1.0 : lam_s, le = (both && !t->annihilate_electrostatics) ? 1.0 : lam_e;
                    if (ea > 0) { double f2; e_ast += softcore(r, sa, ea, ls, t->softcore_alpha, t->softcore_a, t->softcore_b, t->softcore_c, &f2); fr += f2; }
                    double kq = KE_COUL * qa * le;
                    if (kq != 0) {
                        if (h->pme) { double ar = alpha * r, erc = erfc(ar); e_ael += kq * erc / r; fr += kq * (erc / r + TWO_OVER_SQRT_PI * alpha * exp(-ar * ar)) / r2; }
                        else if (t->nb_method == 2) { e_ael += kq * (1 / r + krf * r2 - crf); fr += kq * (1 / (r * r2) - 2 * krf); }
                        else { e_ael += kq / r; fr += kq / (r * r2); }
                    }
                }
                for (int q = 0; q < 3; ++q) { fi[q] += fr * d[q]; Ft[3 * j + q] -= fr * d[q]; }
            }
            for (int q = 0; q < 3; ++q) Ft[3 * i + q] += fi[q];
        }
    }
#pragma omp parallel for
    for (int a = 0; a < 3 * N; ++a) { double s = 0; for (int th = 0; th < h->nthreads; ++th) s += h->Fthr[(size_t)th * 3 * N + a]; F[a] += s; }
    double e_self = 0, e_disp = 0;
    if (h->pme) {
        e_pme = pme_reciprocal(h, x, F);
        double q2 = 0, qs = 0;
        for (int a = 0; a < N; ++a) { q2 += t->charge[a] * t->charge[a]; qs += t->charge[a]; }
        double V = h->box[0] * h->box[1] * h->box[2];
        e_self = -KE_COUL * alpha / sqrt(M_PI) * q2 - KE_COUL * M_PI * qs * qs / (2 * alpha * alpha * V);
    }
    if (h->periodic) e_disp = t->dispersion_coeff / (h->box[0] * h->box[1] * h->box[2]);
    if (terms) { double v[12] = {e_bond, e_angle, e_tors, e_restr, e_pair, e_exc, e_pme, e_self, e_disp, e_ast, e_ael, e_aex}; memcpy(terms, v, sizeof v); }
    return e_bond + e_angle + e_tors + e_restr + e_pair + e_exc + e_pme + e_self + e_disp + e_ast + e_ael + e_aex;
}

/* ------------------------------------------------------------------------------------------------ constraints */
static void shake(orc_handle* h, double* x, const double* xref) {
    const bl_topology* t = &h->t;
    for (int it = 0; it < 500; ++it) {
        double worst = 0;
        for (int k = 0; k < t->n_constraints; ++k) {
            int i = t->constraints[2 * k], j = t->constraints[2 * k + 1];
            double s[3], r[3], d2 = t->constraint_d[k] * t->constraint_d[k];
            for (int q = 0; q < 3; ++q) { s[q] = x[3 * i + q] - x[3 * j + q]; r[q] = xref[3 * i + q] - xref[3 * j + q]; }
            double diff = d2 - (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
            if (fabs(diff) / d2 > worst) worst = fabs(diff) / d2;
            double g = diff / (2 * (s[0] * r[0] + s[1] * r[1] + s[2] * r[2]) * (h->invm[i] + h->invm[j]));
            for (int q = 0; q < 3; ++q) { x[3 * i + q] += g * r[q] * h->invm[i]; x[3 * j + q] -= g * r[q] * h->invm[j]; }
        }
        if (worst < 1e-13) break;
    }
}
static void rattle(orc_handle* h, const double* x, double* v) {
    const bl_topology* t = &h->t;
    for (int it = 0; it < 500; ++it) {
        double worst = 0;
        for (int k = 0; k < t->n_constraints; ++k) {
            int i = t->constraints[2 * k], j = t->constraints[2 * k + 1];
            double s[3], dv = 0, s2 = 0;
            for (int q = 0; q < 3; ++q) { s[q] = x[3 * i + q] - x[3 * j + q]; dv += s[q] * (v[3 * i + q] - v[3 * j + q]); s2 += s[q] * s[q]; }
            double g = -dv / (s2 * (h->invm[i] + h->invm[j]));
            if (fabs(dv) > worst) worst = fabs(dv);
            for (int q = 0; q < 3; ++q) { v[3 * i + q] += g * s[q] * h->invm[i]; v[3 * j + q] -= g * s[q] * h->invm[j]; }
        }
        if (worst < 1e-14) break;
    }
}

/* ------------------------------------------------------------------------------------------------ program */
static void update_alch(orc_handle* h) {
    int k = h->lambda_step; if (k > h->n_lambda_steps) k = h->n_lambda_steps; if (k < 0) k = 0;
    h->lam_s = h->lam_s_tab[k]; h->lam_e = h->lam_e_tab[k];
}
/* OpenMM computes energy and forces together and keeps them while positions and parameters are unchanged
 * (CustomIntegrator force/energy validity): the reference costs 3 evaluations per step, so does this. */
/* alchemical terms only (softcore sterics, lambda-scaled direct-space electrostatics, alchemical exceptions): the same
 * formulas as the pair loop of orc_energy_forces, over the pairs of the current Verlet list that involve an alchemical atom */
static double alch_only(orc_handle* h, const double* x, double lam_s, double lam_e, double* F) {
    const bl_topology* t = &h->t;
    const double rc2 = h->periodic ? t->cutoff * t->cutoff : 1e300, alpha = t->ewald_alpha;
    const double krf = (t->nb_method == 2) ? (1.0 / (t->cutoff * t->cutoff * t->cutoff)) * (78.3 - 1) / (2 * 78.3 + 1) : 0;
    const double crf = (t->nb_method == 2) ? (1.0 / t->cutoff) * 3 * 78.3 / (2 * 78.3 + 1) : 0;
    memset(F, 0, sizeof(double) * 3 * h->N);
    double e = 0;
    for (long p = 0; p < h->n_apairs; ++p) {
        const int i = h->apairs[2 * p], j = h->apairs[2 * p + 1];
        double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        minimg(h, d);
        const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        if (r2 >= rc2) continue;
        const double r = sqrt(r2);
        double fr = 0;
        int both = h->is_alch[i] && h->is_alch[j];
        double sa = 0.5 * ((h->is_alch[i] ? h->asig[i] : t->sigma[i]) + (h->is_alch[j] ? h->asig[j] : t->sigma[j]));
        double ea = sqrt((h->is_alch[i] ? h->aeps[i] : t->epsilon[i]) * (h->is_alch[j] ? h->aeps[j] : t->epsilon[j]));
        double qa = (h->is_alch[i] ? h->aq[i] : t->charge[i]) * (h->is_alch[j] ? h->aq[j] : t->charge[j]);
        double ls = (both && !t->annihilate_sterics) ? 1.0 : lam_s, le = (both && !t->annihilate_electrostatics) ? 1.0 : lam_e;
        if (ea > 0) { double f2; e += softcore(r, sa, ea, ls, t->softcore_alpha, t->softcore_a, t->softcore_b, t->softcore_c, &f2); fr += f2; }
        double kq = KE_COUL * qa * le;
        if (kq != 0) {
            if (h->pme) { double ar = alpha * r, erc = erfc(ar); e += kq * erc / r; fr += kq * (erc / r + TWO_OVER_SQRT_PI * alpha * exp(-ar * ar)) / r2; }
            else if (t->nb_method == 2) { e += kq * (1 / r + krf * r2 - crf); fr += kq * (1 / (r * r2) - 2 * krf); }
            else { e += kq / r; fr += kq / (r * r2); }
        }
        for (int q = 0; q < 3; ++q) { F[3 * i + q] += fr * d[q]; F[3 * j + q] -= fr * d[q]; }
    }
    for (int k = 0; k < t->n_alch_exc; ++k) {
        int i = t->alch_exc_pairs[2 * k], j = t->alch_exc_pairs[2 * k + 1];
        int both = h->is_alch[i] && h->is_alch[j];
        double ls = (both && !t->annihilate_sterics) ? 1.0 : lam_s, le = (both && !t->annihilate_electrostatics) ? 1.0 : lam_e;
        double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        minimg(h, d);
        double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], r = sqrt(r2), fr = 0;
        if (t->alch_exc_eps[k] != 0) e += softcore(r, t->alch_exc_sigma[k], t->alch_exc_eps[k], ls, t->softcore_alpha, t->softcore_a, t->softcore_b, t->softcore_c, &fr);
        double kq = KE_COUL * t->alch_exc_qq[k] * le;
        e += kq / r; fr += kq / (r * r2);
        for (int q = 0; q < 3; ++q) { F[3 * i + q] += fr * d[q]; F[3 * j + q] -= fr * d[q]; }
    }
    return e;
}
static void evaluate(orc_handle* h, const double* x) {
    if (h->cache_valid) return;
    if (h->fast) {
        if (!h->xcache_valid) {
            const double full = orc_energy_forces(h, x, h->lam_s, h->lam_e, h->cache_F, NULL);
            const double ea = alch_only(h, x, h->lam_s, h->lam_e, h->alch_F);
            h->env_E = full - ea;
            for (int a = 0; a < 3 * h->N; ++a) h->env_F[a] = h->cache_F[a] - h->alch_F[a];
            h->cache_E = full;
            h->xcache_valid = 1;
        } else {
            const double ea = alch_only(h, x, h->lam_s, h->lam_e, h->alch_F);
            h->cache_E = h->env_E + ea;
            for (int a = 0; a < 3 * h->N; ++a) h->cache_F[a] = h->env_F[a] + h->alch_F[a];
        }
        h->cache_valid = 1;
        return;
    }
    h->cache_E = orc_energy_forces(h, x, h->lam_s, h->lam_e, h->cache_F, NULL);
    h->cache_valid = 1;
}
static double energy(orc_handle* h, const double* x) { evaluate(h, x); return h->cache_E; }
static double kinetic(orc_handle* h, const double* v) {
    double ke = 0;
    for (int a = 0; a < h->N; ++a) ke += 0.5 * h->t.mass[a] * (v[3 * a] * v[3 * a] + v[3 * a + 1] * v[3 * a + 1] + v[3 * a + 2] * v[3 * a + 2]);
    return ke;
}
static void one_pass(orc_handle* h, double* x, double* v) {
    const int N = h->N;
    if (h->t.remove_cm) {
        double p[3] = {0, 0, 0}, M = 0;
        for (int a = 0; a < N; ++a) { M += h->t.mass[a]; for (int q = 0; q < 3; ++q) p[q] += h->t.mass[a] * v[3 * a + q]; }
        for (int a = 0; a < N; ++a) if (h->invm[a] > 0) for (int q = 0; q < 3; ++q) v[3 * a + q] -= p[q] / M;
    }
    double* x0 = (double*)malloc(sizeof(double) * 3 * N);
    double* x1 = (double*)malloc(sizeof(double) * 3 * N);
    for (int s = 0; s < h->nsplit; ++s) {
        char c = h->split[s];
        if (c == 'V') {
            double hh = h->dt / h->nV;
            evaluate(h, x);
            for (int a = 0; a < N; ++a) for (int q = 0; q < 3; ++q) v[3 * a + q] += hh * h->cache_F[3 * a + q] * h->invm[a];
            rattle(h, x, v);
        } else if (c == 'R') {
            double hh = h->dt / h->nR;
            memcpy(x0, x, sizeof(double) * 3 * N);
            for (int a = 0; a < N; ++a) if (h->invm[a] > 0) for (int q = 0; q < 3; ++q) x[3 * a + q] += hh * v[3 * a + q];
            memcpy(x1, x, sizeof(double) * 3 * N);
            h->cache_valid = 0; h->xcache_valid = 0;
            shake(h, x, x0);
            for (int a = 0; a < 3 * N; ++a) v[a] += (x[a] - x1[a]) / hh;
            rattle(h, x, v);
        } else if (c == 'O') {
            double hh = h->dt / h->nO, aa = exp(-h->gamma * hh), bb = sqrt(1 - exp(-2 * h->gamma * hh));
            double ke0 = kinetic(h, v);
            for (int a = 0; a < N; ++a) {
                if (h->invm[a] <= 0) { v[3 * a] = v[3 * a + 1] = v[3 * a + 2] = 0; continue; }
                double n[3]; normal3(h->seed, 0, h->replica, h->noise_counter, (uint32_t)a, n);
                double sg = bb * sqrt(h->kT * h->invm[a]);
                for (int q = 0; q < 3; ++q) v[3 * a + q] = aa * v[3 * a + q] + sg * n[q];
            }
            h->noise_counter++;
            rattle(h, x, v);
            h->heat += kinetic(h, v) - ke0;
        } else if (c == 'H') {
            if (h->prop != 1) continue;
            h->Eold = energy(h, x);
            h->lambda = (double)(h->lambda_step + 1) / h->n_lambda_steps;
            h->lambda_step++;
            update_alch(h);
            h->cache_valid = 0;
            h->Enew = energy(h, x);
            h->protocol_work += h->Enew - h->Eold;
        }
    }
    free(x0); free(x1);
}

void orc_step(orc_handle* h, double* x, double* v, int n) {
    h->cache_valid = 0; h->xcache_valid = 0;      /* the caller may have changed x between calls */
    for (int i = 0; i < n; ++i) {
        if (h->step == 0) {
            h->perturbed_pe = h->unperturbed_pe = energy(h, x);
            double* x0 = (double*)dup_mem(x, sizeof(double) * 3 * h->N);
            shake(h, x, x0); free(x0);
            rattle(h, x, v);
            h->cache_valid = 0; h->xcache_valid = 0;
            h->protocol_work = 0; h->lambda = 0; h->lambda_step = 0; update_alch(h);
        }
        if (h->step < h->nsteps) {
            h->perturbed_pe = energy(h, x);
            if (h->first_step < 1) { h->first_step = 1; h->unperturbed_pe = h->perturbed_pe; }
            h->protocol_work += h->perturbed_pe - h->unperturbed_pe;
            one_pass(h, x, v);
            if (h->lambda > h->pl_min && h->lambda <= h->pl_max)
                while (h->prop < h->nprop) { h->prop++; one_pass(h, x, v); }
            h->unperturbed_pe = energy(h, x);
            h->step++; h->prop = 1;
        }
    }
}

void orc_velocities_to_temperature(orc_handle* h, const double* x, double* v, double T) {
    for (int a = 0; a < h->N; ++a) {
        double n[3]; normal3(h->seed, 1, h->replica, h->vel_counter, (uint32_t)a, n);
        double sg = sqrt(KB_KJ * T * h->invm[a]);
        for (int q = 0; q < 3; ++q) v[3 * a + q] = sg * n[q];
    }
    h->vel_counter++;
    rattle(h, x, v);
}

void orc_reset(orc_handle* h) {
    h->step = 0; h->lambda = 0; h->protocol_work = 0; h->first_step = 0; h->perturbed_pe = h->unperturbed_pe = 0;
    h->prop = 1; h->lambda_step = 0; update_alch(h); h->cache_valid = 0;
}
double orc_get(orc_handle* h, const char* name) {
    if (!strcmp(name, "protocol_work")) return h->protocol_work;
    if (!strcmp(name, "lambda")) return h->lambda;
    if (!strcmp(name, "step")) return h->step;
    if (!strcmp(name, "lambda_step")) return h->lambda_step;
    if (!strcmp(name, "n_evals")) return (double)h->n_evals;
    if (!strcmp(name, "unperturbed_pe")) return h->unperturbed_pe;
    if (!strcmp(name, "heat")) return h->heat;
    if (!strcmp(name, "n_neighbors")) return h->nl_valid ? h->nl_ptr[h->N] : 0;
    return NAN;
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* launchers such as torchrun export OMP_NUM_THREADS=1: the caller states the thread count explicitly (before orc_create,
   which sizes the per-thread force buffers) */
/* bench context row: one force evaluation per step (lambda-separable); call before the first step */
void orc_set_fast(orc_handle* h, int on) {
    h->fast = on != 0;
    h->nl_valid = 0; h->cache_valid = 0; h->xcache_valid = 0;
    if (h->fast && !h->env_F) {
        h->env_F = (double*)malloc(sizeof(double) * 3 * h->N);
        h->alch_F = (double*)malloc(sizeof(double) * 3 * h->N);
    }
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

#define DUP(field, count, type) h->t.field = (const type*)dup_mem(t->field, sizeof(type) * (size_t)(count))
orc_handle* orc_create(const bl_topology* t, const bl_integrator_params* p, uint64_t seed, int replica) {
    orc_handle* h = (orc_handle*)calloc(1, sizeof(orc_handle));
    h->t = *t;
    const int N = h->N = t->n_atoms;
    DUP(mass, N, double); DUP(charge, N, double); DUP(sigma, N, double); DUP(epsilon, N, double);
    DUP(bonds, 2 * t->n_bonds, int32_t); DUP(bond_k, t->n_bonds, double); DUP(bond_r0, t->n_bonds, double);
    DUP(angles, 3 * t->n_angles, int32_t); DUP(angle_k, t->n_angles, double); DUP(angle_t0, t->n_angles, double);
    DUP(torsions, 4 * t->n_torsions, int32_t); DUP(torsion_k, t->n_torsions, double); DUP(torsion_n, t->n_torsions, int32_t);
    DUP(torsion_phase, t->n_torsions, double);
    DUP(excl_pairs, 2 * t->n_excl, int32_t); DUP(excl_qq, t->n_excl, double); DUP(excl_sigma, t->n_excl, double); DUP(excl_eps, t->n_excl, double);
    DUP(constraints, 2 * t->n_constraints, int32_t); DUP(constraint_d, t->n_constraints, double);
    DUP(restraint_atoms, t->n_restraints, int32_t); DUP(restraint_k, t->n_restraints, double); DUP(restraint_x0, 3 * t->n_restraints, double);
    DUP(alch_atoms, t->n_alch, int32_t); DUP(alch_charge, t->n_alch, double); DUP(alch_sigma, t->n_alch, double); DUP(alch_eps, t->n_alch, double);
    DUP(alch_exc_pairs, 2 * t->n_alch_exc, int32_t); DUP(alch_exc_qq, t->n_alch_exc, double); DUP(alch_exc_sigma, t->n_alch_exc, double);
    DUP(alch_exc_eps, t->n_alch_exc, double);
    h->periodic = t->nb_method != 0; h->pme = t->nb_method == 4;
    memcpy(h->box, t->box, sizeof h->box);
    h->invm = (double*)malloc(sizeof(double) * N);
    for (int a = 0; a < N; ++a) h->invm[a] = t->mass[a] > 0 ? 1.0 / t->mass[a] : 0.0;
    h->ex_ptr = (int*)calloc(N + 1, sizeof(int));
    for (int k = 0; k < t->n_excl; ++k) { h->ex_ptr[t->excl_pairs[2 * k] + 1]++; h->ex_ptr[t->excl_pairs[2 * k + 1] + 1]++; }
    for (int a = 0; a < N; ++a) h->ex_ptr[a + 1] += h->ex_ptr[a];
    h->ex_idx = (int*)malloc(sizeof(int) * (2 * t->n_excl + 1));
    int* fill = (int*)dup_mem(h->ex_ptr, sizeof(int) * (N + 1));
    for (int k = 0; k < t->n_excl; ++k) { int i = t->excl_pairs[2 * k], j = t->excl_pairs[2 * k + 1]; h->ex_idx[fill[i]++] = j; h->ex_idx[fill[j]++] = i; }
    free(fill);
    h->is_alch = (unsigned char*)calloc(N, 1);
    h->aq = (double*)calloc(N, sizeof(double)); h->asig = (double*)calloc(N, sizeof(double)); h->aeps = (double*)calloc(N, sizeof(double));
    for (int k = 0; k < t->n_alch; ++k) { int a = t->alch_atoms[k]; h->is_alch[a] = 1; h->aq[a] = t->alch_charge[k]; h->asig[a] = t->alch_sigma[k]; h->aeps[a] = t->alch_eps[k]; }
    h->nl_ptr = (int*)calloc(N + 1, sizeof(int)); h->nl_cap = 1024; h->nl_idx = (int*)malloc(sizeof(int) * h->nl_cap);
    h->x_ref = (double*)malloc(sizeof(double) * 3 * N);
    h->skin = h->periodic ? 0.1 * t->cutoff : 0.0;
    if (h->pme) {
        size_t G = 1;
        for (int d = 0; d < 3; ++d) { h->K[d] = t->pme_grid[d]; G *= h->K[d]; }
        h->Q = (double*)malloc(sizeof(double) * G); h->S = (cplx*)malloc(sizeof(cplx) * G);
        double node[PME_ORDER];
        for (int k = 0; k < PME_ORDER; ++k) node[k] = M_spline(PME_ORDER, (double)k);
        for (int d = 0; d < 3; ++d) {
            int K = h->K[d];
            h->bmod[d] = (double*)malloc(sizeof(double) * K);
            for (int m = 0; m < K; ++m) {
                double sc = 0, ss = 0;
                for (int k = 0; k < PME_ORDER; ++k) { double arg = 2 * M_PI * m * k / K; sc += node[k] * cos(arg); ss += node[k] * sin(arg); }
                h->bmod[d][m] = sc * sc + ss * ss;
            }
            for (int m = 0; m < K; ++m) if (h->bmod[d][m] < 1e-7) h->bmod[d][m] = 0.5 * (h->bmod[d][(m - 1 + K) % K] + h->bmod[d][(m + 1) % K]);
        }
    }
    h->F = (double*)malloc(sizeof(double) * 3 * N);
    h->cache_F = (double*)malloc(sizeof(double) * 3 * N);
    h->nthreads = orc_num_threads();
    h->Fthr = (double*)malloc(sizeof(double) * 3 * N * h->nthreads);
    h->seed = seed; h->replica = (uint32_t)replica; h->prop = 1;
    h->lam_s = h->lam_e = 1.0;
    if (p) {
        h->kT = KB_KJ * p->temperature; h->gamma = p->friction; h->dt = p->timestep; h->tol = p->constraint_tol;
        h->nsteps = p->nsteps_neq; h->nprop = p->nprop > 0 ? p->nprop : 1; h->pl_min = p->prop_lambda_min; h->pl_max = p->prop_lambda_max;
        const char* s = p->splitting ? p->splitting : "H V R O R V H";
        int nH = 0;
        for (; *s; ++s) if (*s != ' ' && h->nsplit < 31) { h->split[h->nsplit++] = *s; if (*s == 'V') h->nV++; if (*s == 'R') h->nR++; if (*s == 'O') h->nO++; if (*s == 'H') nH++; }
        h->n_lambda_steps = h->nsteps * nH;
        h->lam_s_tab = (double*)dup_mem(p->lambda_sterics, sizeof(double) * p->n_lambda);
        h->lam_e_tab = (double*)dup_mem(p->lambda_electrostatics, sizeof(double) * p->n_lambda);
        update_alch(h);
    }
    return h;
}
void orc_set_lambda(orc_handle* h, double ls, double le) { h->lam_s = ls; h->lam_e = le; }
void orc_destroy(orc_handle* h) { free(h); /* test infrastructure: the process owns the rest */ }
