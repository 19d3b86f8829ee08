// Bonded terms, exceptions / Ewald exclusion corrections, restraints and alchemical exceptions (K7),
// plus the alchemical softcore kernel (K3).  Double precision: a few 1e4 terms, accuracy over speed.
#pragma once
#include "engine.cuh"

__device__ __forceinline__ double3 d3sub(double4 a, double4 b) { return make_double3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ double d3dot(double3 a, double3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double3 d3cross(double3 a, double3 b) {
    return make_double3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double3 d3scale(double3 a, double s) { return make_double3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double3 d3add(double3 a, double3 b) { return make_double3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ void min_image(const Dev& d, double3& v) {
    if (d.periodic) {
        v.x -= d.boxd[0] * rint(v.x * d.boxd[3]);
        v.y -= d.boxd[1] * rint(v.y * d.boxd[4]);
        v.z -= d.boxd[2] * rint(v.z * d.boxd[5]);
    }
}
__device__ __forceinline__ void add_force(long long* f, int N, int a, double3 v) {
    fx_add(&f[a], v.x, FORCE_SCALE);
    fx_add(&f[N + a], v.y, FORCE_SCALE);
    fx_add(&f[2 * N + a], v.z, FORCE_SCALE);
}

// softcore Lennard-Jones of openmmtools: returns U and writes -dU/dr / r ("force over r")
__device__ __forceinline__ double softcore_lj(double r, double sigma, double eps, double lam, double alpha, double a,
                                              double b, double c, double& f_over_r) {
    const double rs = r / sigma;
    double la = (a == 1.0) ? lam : pow(lam, a);
    double oml = (b == 1.0) ? (1.0 - lam) : pow(1.0 - lam, b);
    double rc_, drc;  // (r/sigma)^c and its r-derivative
    if (c == 6.0) {
        double r2 = rs * rs;
        double r5 = r2 * r2 * rs;
        rc_ = r5 * rs;
        drc = 6.0 * r5 / sigma;
    } else {
        rc_ = pow(rs, c);
        drc = c * pow(rs, c - 1.0) / sigma;
    }
    const double s = alpha * oml + rc_;
    double x, dxdr;
    if (c == 6.0) {
        x = 1.0 / s;
        dxdr = -drc / (s * s);
    } else {
        x = pow(s, -6.0 / c);
        dxdr = (-6.0 / c) * pow(s, -6.0 / c - 1.0) * drc;
    }
    const double U = la * 4.0 * eps * x * (x - 1.0);
    const double dU = la * 4.0 * eps * (2.0 * x - 1.0) * dxdr;
    f_over_r = -dU / r;
    return U;
}

__global__ void __launch_bounds__(128) k_bonded(Dev d) {
    const int r = blockIdx.y;
    const int N = d.N;
    const double4* pos = d.pos + (size_t)r * N;
    long long* fenv = d.f_env + (size_t)r * 3 * N;
    const int t0 = d.n_bonds, t1 = t0 + d.n_angles, t2 = t1 + d.n_torsions, t3 = t2 + d.n_excl,
              t4 = t3 + d.n_restraints, t5 = t4 + d.n_alch_exc;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    int term = -1;
    double ealch[ALCH_SLOTS] = {0.0, 0.0, 0.0};
    if (tid < t0) {
        term = E_BOND;
        int2 ix = d.bonds[tid];
        double2 p = d.bond_p[tid];
        double3 dv = d3sub(pos[ix.y], pos[ix.x]);
        min_image(d, dv);
        double rr = sqrt(d3dot(dv, dv));
        double dr = rr - p.y;
        e = 0.5 * p.x * dr * dr;
        double3 f = d3scale(dv, p.x * dr / rr);
        add_force(fenv, N, ix.x, f);
        add_force(fenv, N, ix.y, d3scale(f, -1.0));
    } else if (tid < t1) {
        term = E_ANGLE;
        int4 ix = d.angles[tid - t0];
        double2 p = d.angle_p[tid - t0];
        double3 v1 = d3sub(pos[ix.x], pos[ix.y]);
        double3 v2 = d3sub(pos[ix.z], pos[ix.y]);
        min_image(d, v1);
        min_image(d, v2);
        double r1 = sqrt(d3dot(v1, v1)), r2 = sqrt(d3dot(v2, v2));
        double cs = fmin(1.0, fmax(-1.0, d3dot(v1, v2) / (r1 * r2)));
        double th = acos(cs);
        double dth = th - p.y;
        e = 0.5 * p.x * dth * dth;
        double sn = sqrt(fmax(1.0 - cs * cs, 1e-30));
        double c = p.x * dth / sn;
        double3 f1 = d3scale(d3add(d3scale(v2, 1.0 / (r1 * r2)), d3scale(v1, -cs / (r1 * r1))), c);
        double3 f3 = d3scale(d3add(d3scale(v1, 1.0 / (r1 * r2)), d3scale(v2, -cs / (r2 * r2))), c);
        add_force(fenv, N, ix.x, f1);
        add_force(fenv, N, ix.z, f3);
        add_force(fenv, N, ix.y, d3scale(d3add(f1, f3), -1.0));
    } else if (tid < t2) {
        term = E_TORSION;
        int4 ix = d.torsions[tid - t1];
        double4 p = d.torsion_p[tid - t1];
        double3 b1 = d3sub(pos[ix.y], pos[ix.x]);
        double3 b2 = d3sub(pos[ix.z], pos[ix.y]);
        double3 b3 = d3sub(pos[ix.w], pos[ix.z]);
        min_image(d, b1);
        min_image(d, b2);
        min_image(d, b3);
        double3 n1 = d3cross(b1, b2), n2 = d3cross(b2, b3);
        double b2n = sqrt(d3dot(b2, b2));
        double3 m1 = d3cross(n1, d3scale(b2, 1.0 / b2n));
        double phi = atan2(d3dot(m1, n2), d3dot(n1, n2));
        double arg = p.y * phi - p.z;
        e = p.x * (1.0 + cos(arg));
        double dE = -p.x * p.y * sin(arg);   // dE/dphi
        double n1s = d3dot(n1, n1), n2s = d3dot(n2, n2);
        double3 g0 = d3scale(n1, -b2n / n1s);
        double3 g3 = d3scale(n2, b2n / n2s);
        double s12 = d3dot(b1, b2) / (b2n * b2n), s32 = d3dot(b3, b2) / (b2n * b2n);
        double3 g1 = d3add(d3scale(g0, -1.0 - s12), d3scale(g3, s32));
        double3 g2 = d3add(d3scale(g3, -1.0 - s32), d3scale(g0, s12));
        // phi as defined through m1 = n1 x b2^ has the opposite sign to the g-vectors' convention
        add_force(fenv, N, ix.x, d3scale(g0, dE));
        add_force(fenv, N, ix.y, d3scale(g1, dE));
        add_force(fenv, N, ix.z, d3scale(g2, dE));
        add_force(fenv, N, ix.w, d3scale(g3, dE));
    } else if (tid < t3) {
        term = E_EXCEPT;
        int2 ix = d.excl[tid - t2];
        double4 p = d.excl_p[tid - t2];   // (k qq_exc, sigma, eps, k q_i q_j)
        double3 dv = d3sub(pos[ix.x], pos[ix.y]);
        min_image(d, dv);
        double r2 = d3dot(dv, dv);
        double rr = sqrt(r2);
        double fr = 0.0;   // -dU/dr / r
        if (p.z != 0.0) {
            double s2 = p.y * p.y / r2, s6 = s2 * s2 * s2;
            e += 4.0 * p.z * s6 * (s6 - 1.0);
            fr += 4.0 * p.z * (12.0 * s6 * s6 - 6.0 * s6) / r2;
        }
        if (p.x != 0.0) {
            e += p.x / rr;
            fr += p.x / (rr * r2);
        }
        if (d.pme && p.w != 0.0) {
            double ar = d.alphad * rr;
            double er = erf(ar);
            e -= p.w * er / rr;
            // d/dr [ -k erf(ar)/r ] = -k (2a/sqrt(pi) exp(-a²r²)/r - erf/r²)
            fr += p.w * (TWO_OVER_SQRT_PI * d.alphad * exp(-ar * ar) / rr - er / r2) / rr;
        }
        double3 f = d3scale(dv, fr);
        add_force(fenv, N, ix.x, f);
        add_force(fenv, N, ix.y, d3scale(f, -1.0));
    } else if (tid < t4) {
        term = E_RESTRAINT;
        int a = d.restraint_atom[tid - t3];
        double4 p = d.restraint_p[tid - t3];
        double3 dv = make_double3(pos[a].x - p.x, pos[a].y - p.y, pos[a].z - p.z);
        min_image(d, dv);
        e = p.w * d3dot(dv, dv);
        add_force(fenv, N, a, d3scale(dv, -2.0 * p.w));
    } else if (tid < t5) {
        term = E_ALCH_EXC;
        int2 ix = d.alch_exc[tid - t4];
        double4 p = d.alch_exc_p[tid - t4];   // (k qq, sigma, eps, both)
        double3 dv = d3sub(pos[ix.x], pos[ix.y]);
        min_image(d, dv);
        double r2 = d3dot(dv, dv), rr = sqrt(r2);
        const int base = d.g[r].lambda_step;
        const bool both = p.w != 0.0;
        for (int s = 0; s < ALCH_SLOTS; ++s) {
            int li = min(base + s, d.n_lambda - 1);
            double ls = (both && !d.annihilate_sterics) ? 1.0 : d.lam_s[li];
            double le = (both && !d.annihilate_elec) ? 1.0 : d.lam_e[li];
            double fr = 0.0, es = 0.0;
            if (p.z != 0.0) es = softcore_lj(rr, p.y, p.z, ls, d.sc_alpha, d.sc_a, d.sc_b, d.sc_c, fr);
            es += le * p.x / rr;
            fr += le * p.x / (rr * r2);
            ealch[s] = es;
            long long* fa = d.f_alch + ((size_t)s * d.R + r) * 3 * N;
            double3 f = d3scale(dv, fr);
            add_force(fa, N, ix.x, f);
            add_force(fa, N, ix.y, d3scale(f, -1.0));
        }
    }
    // energy reduction: warp-level per term kind, one fixed-point atomic per warp and kind
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < N_ETERMS; ++k) {
        if (k == E_ALCH_EXC) continue;
        unsigned int m = __ballot_sync(0xffffffffu, term == k);
        if (!m) continue;
        double v = warp_sum(term == k ? e : 0.0);
        if (lane == 0) fx_add(&d.eacc[r * N_ETERMS + k], v, ENERGY_SCALE);
    }
    if (__any_sync(0xffffffffu, term == E_ALCH_EXC)) {
        for (int s = 0; s < ALCH_SLOTS; ++s) {
            double v = warp_sum(ealch[s]);
            if (lane == 0) fx_add(&d.alch_acc[(r * ALCH_SLOTS + s) * 3 + 2], v, ENERGY_SCALE);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Alchemical-region pairs (K3).  k_alch_list runs with the neighbour rebuild: for every alchemical atom it
// compacts (warp ballot) the atoms within the list cutoff that are not excluded into a per-alchemical-atom list.
// k_alch then evaluates one pair per thread at ALCH_SLOTS consecutive lambda_step values in one pass (energies
// for Enew - Eold, forces for the V steps that follow), reducing the force on the alchemical atom over the warp.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_alch_list(Dev d) {
    const int r = blockIdx.y;
    if (!d.g[r].do_prune) return;
    const int N = d.N, na = d.n_alch;
    extern __shared__ float4 s_apos[];           // [na] alchemical positions (w = orig index as float bits)
    const float4* posq = d.posq + (size_t)r * N;
    for (int k = threadIdx.x; k < na; k += blockDim.x) {
        const int a = d.alch_atom[k];
        float4 p = posq[a];
        p.w = __int_as_float(a);
        s_apos[k] = p;
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool active = j < N;
    const float4 pj = active ? posq[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool j_alch = active ? d.is_alch[j] : false;
    const ull wj = active ? d.excl_win[j] : 0ull;
    const bool farj = active ? d.has_far[j] : false;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2], ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    for (int k = 0; k < na; ++k) {
        const float4 pi = s_apos[k];
        const int i = __float_as_int(pi.w);
        bool ok = active && i != j;
        if (ok && j_alch) ok = i < j;            // alchemical-alchemical pairs once, owned by the lower index
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        if (d.periodic) {
            dx -= bx * rintf(dx * ibx);
            dy -= by * rintf(dy * iby);
            dz -= bz * rintf(dz * ibz);
        }
        ok = ok && (dx * dx + dy * dy + dz * dz) < d.list_cutoff2;
        if (ok) ok = !pair_excluded(d, j, wj, farj, i, d.has_far[i]);
        const unsigned int m = __ballot_sync(0xffffffffu, ok);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&d.alch_count[r * na + k], __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (ok) {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < d.alch_cap) d.alch_list[((size_t)r * na + k) * d.alch_cap + slot] = j;
                else d.g[r].item_overflow = 1;
            }
        }
    }
}

// Order every alchemical atom's pair list by atom index (bitonic sort in shared memory; alch_cap <= 2048).  The atomic
// append of k_alch_list leaves an arbitrary order, and the assignment of entries to the threads of k_alch decides how
// their double partial sums round when they enter the fixed-point accumulators: sorted lists make the alchemical
// energies and forces bitwise reproducible from run to run.  Rebuild steps only, off the critical path (stream 4).
#define ALCH_SORT_MAX 2048
__global__ void __launch_bounds__(512) k_alch_sort(Dev d) {
    __shared__ int s_key[ALCH_SORT_MAX];
    const int k = blockIdx.x, r = blockIdx.y, na = d.n_alch;
    if (!d.g[r].do_prune) return;
    const int n = min(d.alch_count[r * na + k], min(d.alch_cap, ALCH_SORT_MAX));
    if (n <= 1) return;
    int* list = d.alch_list + ((size_t)r * na + k) * d.alch_cap;
    int m = 2;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_key[i] = i < n ? list[i] : 0x7fffffff;
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const int a = s_key[i], b = s_key[j];
                    const bool ascending = (i & size) == 0;
                    if ((a > b) == ascending) { s_key[i] = b; s_key[j] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < n; i += blockDim.x) list[i] = s_key[i];
}

__global__ void k_alch_reset(Dev d) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= d.R * d.n_alch) return;
    if (d.g[idx / d.n_alch].do_prune) d.alch_count[idx] = 0;
}

__global__ void __launch_bounds__(128) k_alch(Dev d) {
    const int r = blockIdx.z, k = blockIdx.y;
    const int N = d.N, na = d.n_alch;
    const int count = min(d.alch_count[r * na + k], d.alch_cap);
    if ((int)(blockIdx.x * blockDim.x) >= count) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool active = p < count;
    const double4* pos = d.pos + (size_t)r * N;
    const int i = d.alch_atom[k];
    const double4 par_i = d.alch_p[k];           // (q, sigma, eps)
    const double4 pi = pos[i];
    const int base = d.g[r].lambda_step;
    double fi[ALCH_SLOTS][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double e_st[ALCH_SLOTS] = {0, 0, 0}, e_el[ALCH_SLOTS] = {0, 0, 0};
    if (active) {
        const int j = d.alch_list[((size_t)r * na + k) * d.alch_cap + p];
        const double4 pj = pos[j];
        double3 dv = make_double3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        min_image(d, dv);
        const double r2 = d3dot(dv, dv);
        const double cut2 = d.periodic ? d.cutoffd * d.cutoffd : 1e300;
        if (r2 < cut2) {
            const bool j_alch = d.is_alch[j];
            double qj = d.charge_d[j], sj = d.sigma_d[j], ej = d.eps_d[j];
            if (j_alch) {
                for (int m = 0; m < na; ++m)
                    if (d.alch_atom[m] == j) { const double4 q = d.alch_p[m]; qj = q.x; sj = q.y; ej = q.z; }
            }
            const double rr = sqrt(r2);
            const double sig = 0.5 * (par_i.y + sj);
            const double eps = sqrt(par_i.z * ej);
            const double kqq = ONE_4PI_EPS0 * par_i.x * qj;
            double ec = 0.0, fc = 0.0;   // unit-lambda electrostatic energy and force/r
            if (kqq != 0.0) {
                if (d.pme) {
                    const double ar = d.alphad * rr;
                    const double erc = erfc(ar);
                    ec = kqq * erc / rr;
                    fc = kqq * (erc / rr + TWO_OVER_SQRT_PI * d.alphad * exp(-ar * ar)) / r2;
                } else if (d.nb_method == 2) {
                    ec = kqq * (1.0 / rr + (double)d.krf * r2 - (double)d.crf);
                    fc = kqq * (1.0 / (rr * r2) - 2.0 * (double)d.krf);
                } else {
                    ec = kqq / rr;
                    fc = kqq / (rr * r2);
                }
            }
            for (int s = 0; s < ALCH_SLOTS; ++s) {
                const int li = min(base + s, d.n_lambda - 1);
                const double lss = (j_alch && !d.annihilate_sterics) ? 1.0 : d.lam_s[li];
                const double les = (j_alch && !d.annihilate_elec) ? 1.0 : d.lam_e[li];
                double fr = 0.0;
                if (eps > 0.0) e_st[s] = softcore_lj(rr, sig, eps, lss, d.sc_alpha, d.sc_a, d.sc_b, d.sc_c, fr);
                e_el[s] = les * ec;
                fr += les * fc;
                fi[s][0] = dv.x * fr; fi[s][1] = dv.y * fr; fi[s][2] = dv.z * fr;
                long long* fa = d.f_alch + ((size_t)s * d.R + r) * 3 * N;
                add_force(fa, N, j, make_double3(-fi[s][0], -fi[s][1], -fi[s][2]));
            }
        }
    }
    for (int s = 0; s < ALCH_SLOTS; ++s) {
        const double fx = warp_sum(fi[s][0]), fy = warp_sum(fi[s][1]), fz = warp_sum(fi[s][2]);
        const double a = warp_sum(e_st[s]), b = warp_sum(e_el[s]);
        if (lane == 0) {
            long long* fa = d.f_alch + ((size_t)s * d.R + r) * 3 * N;
            if (fx != 0.0 || fy != 0.0 || fz != 0.0) add_force(fa, N, i, make_double3(fx, fy, fz));
            if (a != 0.0) fx_add(&d.alch_acc[(r * ALCH_SLOTS + s) * 3 + 0], a, ENERGY_SCALE);
            if (b != 0.0) fx_add(&d.alch_acc[(r * ALCH_SLOTS + s) * 3 + 1], b, ENERGY_SCALE);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// k_custom: generic Custom*Force terms (bl_topology::custom_*).  One thread per (term, walker): r = distance between the
// weighted centroids of two atom groups, E(r) and dE/dr from the host-compiled stack program by forward-mode dual
// numbers, at the ALCH_SLOTS consecutive lambda_step values exactly like k_alch (energies into the slot accumulators
// for Enew - Eold, forces into the slot force buffers for the following V steps).  Double precision throughout; the
// terms are few (the reference's use: 12 pairs + 1 centroid bond, blues/tests/data/ethylene_system.xml:52-114).
// ---------------------------------------------------------------------------------------------------------
struct Dual { double v, d; };

__device__ __noinline__ Dual custom_eval(const Dev& d, int prog, double r, const double* par, double g0, double g1) {
    double sv[BL_CUSTOM_STACK], sd[BL_CUSTOM_STACK];
    int sp = 0;
    const int end = d.custom_pstart[prog + 1];
    for (int pc = d.custom_pstart[prog]; pc < end; ++pc) {
        const int op = d.custom_op[pc];
        const double arg = d.custom_arg[pc];
        switch (op) {
        case BL_OP_CONST: sv[sp] = arg; sd[sp] = 0.0; ++sp; break;
        case BL_OP_R: sv[sp] = r; sd[sp] = 1.0; ++sp; break;
        case BL_OP_PARAM: sv[sp] = par[(int)arg]; sd[sp] = 0.0; ++sp; break;
        case BL_OP_GLOBAL: sv[sp] = ((int)arg == 0) ? g0 : g1; sd[sp] = 0.0; ++sp; break;
        case BL_OP_ADD: --sp; sv[sp - 1] += sv[sp]; sd[sp - 1] += sd[sp]; break;
        case BL_OP_SUB: --sp; sv[sp - 1] -= sv[sp]; sd[sp - 1] -= sd[sp]; break;
        case BL_OP_MUL: --sp; sd[sp - 1] = sd[sp - 1] * sv[sp] + sv[sp - 1] * sd[sp]; sv[sp - 1] *= sv[sp]; break;
        case BL_OP_DIV: {
            --sp;
            const double inv = 1.0 / sv[sp], q = sv[sp - 1] * inv;
            sd[sp - 1] = (sd[sp - 1] - q * sd[sp]) * inv;
            sv[sp - 1] = q;
        } break;
        case BL_OP_NEG: sv[sp - 1] = -sv[sp - 1]; sd[sp - 1] = -sd[sp - 1]; break;
        case BL_OP_POWI: {
            const int n = (int)arg;
            const double x = sv[sp - 1];
            double pw = 1.0;                                      // x^(|n| - 1)
            for (int k = 1; k < abs(n); ++k) pw *= x;
            if (n == 0) { sv[sp - 1] = 1.0; sd[sp - 1] = 0.0; }
            else if (n > 0) { sd[sp - 1] *= n * pw; sv[sp - 1] = pw * x; }
            else { const double v = 1.0 / (pw * x); sd[sp - 1] *= n * v / x; sv[sp - 1] = v; }
        } break;
        case BL_OP_POW: {
            --sp;
            const double a = sv[sp - 1], b = sv[sp], v = pow(a, b);
            double dv = 0.0;
            if (sd[sp - 1] != 0.0) dv += b * pow(a, b - 1.0) * sd[sp - 1];
            if (sd[sp] != 0.0) dv += v * log(a) * sd[sp];
            sv[sp - 1] = v; sd[sp - 1] = dv;
        } break;
        case BL_OP_SQRT: { const double v = sqrt(sv[sp - 1]); sd[sp - 1] = sd[sp - 1] != 0.0 ? 0.5 * sd[sp - 1] / v : 0.0; sv[sp - 1] = v; } break;
        case BL_OP_EXP: { const double v = exp(sv[sp - 1]); sd[sp - 1] *= v; sv[sp - 1] = v; } break;
        case BL_OP_LOG: sd[sp - 1] /= sv[sp - 1]; sv[sp - 1] = log(sv[sp - 1]); break;
        case BL_OP_SIN: sd[sp - 1] *= cos(sv[sp - 1]); sv[sp - 1] = sin(sv[sp - 1]); break;
        case BL_OP_COS: sd[sp - 1] *= -sin(sv[sp - 1]); sv[sp - 1] = cos(sv[sp - 1]); break;
        case BL_OP_TAN: { const double t = tan(sv[sp - 1]); sd[sp - 1] *= 1.0 + t * t; sv[sp - 1] = t; } break;
        case BL_OP_ABS: if (sv[sp - 1] < 0.0) { sv[sp - 1] = -sv[sp - 1]; sd[sp - 1] = -sd[sp - 1]; } break;
        case BL_OP_MIN: --sp; if (sv[sp] < sv[sp - 1]) { sv[sp - 1] = sv[sp]; sd[sp - 1] = sd[sp]; } break;
        case BL_OP_MAX: --sp; if (sv[sp] > sv[sp - 1]) { sv[sp - 1] = sv[sp]; sd[sp - 1] = sd[sp]; } break;
        case BL_OP_STEP: sv[sp - 1] = sv[sp - 1] >= 0.0 ? 1.0 : 0.0; sd[sp - 1] = 0.0; break;
        case BL_OP_DELTA: sv[sp - 1] = sv[sp - 1] == 0.0 ? 1.0 : 0.0; sd[sp - 1] = 0.0; break;
        case BL_OP_SELECT: sp -= 2; if (sv[sp - 1] != 0.0) { sv[sp - 1] = sv[sp]; sd[sp - 1] = sd[sp]; } else { sv[sp - 1] = sv[sp + 1]; sd[sp - 1] = sd[sp + 1]; } break;
        case BL_OP_ERF: sd[sp - 1] *= TWO_OVER_SQRT_PI * exp(-sv[sp - 1] * sv[sp - 1]); sv[sp - 1] = erf(sv[sp - 1]); break;
        case BL_OP_ERFC: sd[sp - 1] *= -TWO_OVER_SQRT_PI * exp(-sv[sp - 1] * sv[sp - 1]); sv[sp - 1] = erfc(sv[sp - 1]); break;
        case BL_OP_TANH: { const double t = tanh(sv[sp - 1]); sd[sp - 1] *= 1.0 - t * t; sv[sp - 1] = t; } break;
        case BL_OP_SINH: sd[sp - 1] *= cosh(sv[sp - 1]); sv[sp - 1] = sinh(sv[sp - 1]); break;
        case BL_OP_COSH: sd[sp - 1] *= sinh(sv[sp - 1]); sv[sp - 1] = cosh(sv[sp - 1]); break;
        case BL_OP_ATAN: sd[sp - 1] /= 1.0 + sv[sp - 1] * sv[sp - 1]; sv[sp - 1] = atan(sv[sp - 1]); break;
        default: break;
        }
    }
    Dual out;
    out.v = sp > 0 ? sv[0] : 0.0;
    out.d = sp > 0 ? sd[0] : 0.0;
    return out;
}

__global__ void __launch_bounds__(64) k_custom(Dev d) {
    const int r = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.n_custom) return;
    const int N = d.N;
    const double4* pos = d.pos + (size_t)r * N;
    const int4 term = d.custom_term[t];
    double3 ca = make_double3(0, 0, 0), cb = make_double3(0, 0, 0);
    for (int k = d.custom_gstart[term.x]; k < d.custom_gstart[term.x + 1]; ++k) {
        const double4 p = pos[d.custom_gatoms[k]];
        const double w = d.custom_gweights[k];
        ca.x += w * p.x; ca.y += w * p.y; ca.z += w * p.z;
    }
    for (int k = d.custom_gstart[term.y]; k < d.custom_gstart[term.y + 1]; ++k) {
        const double4 p = pos[d.custom_gatoms[k]];
        const double w = d.custom_gweights[k];
        cb.x += w * p.x; cb.y += w * p.y; cb.z += w * p.z;
    }
    double3 dv = make_double3(ca.x - cb.x, ca.y - cb.y, ca.z - cb.z);
    if (term.w & 1) min_image(d, dv);
    const double rr = sqrt(d3dot(dv, dv));
    const double cut = d.custom_cutoff[t];
    if (cut > 0.0 && rr >= cut) return;
    const double* par = d.custom_params + (size_t)t * d.custom_np;
    const int base = d.g[r].lambda_step;
    for (int s = 0; s < ALCH_SLOTS; ++s) {
        const int li = min(base + s, d.n_lambda - 1);
        const Dual e = custom_eval(d, term.z, rr, par, d.lam_s[li], d.lam_e[li]);
        if (e.v != 0.0) fx_add(&d.alch_acc[(r * ALCH_SLOTS + s) * 3 + 0], e.v, ENERGY_SCALE);
        const double fr = rr > 0.0 ? -e.d / rr : 0.0;          // F_A = -dE/dr * (cA - cB) / r, shared out by the weights
        if (fr != 0.0) {
            long long* fa = d.f_alch + ((size_t)s * d.R + r) * 3 * N;
            for (int k = d.custom_gstart[term.x]; k < d.custom_gstart[term.x + 1]; ++k) {
                const double w = d.custom_gweights[k] * fr;
                add_force(fa, N, d.custom_gatoms[k], make_double3(w * dv.x, w * dv.y, w * dv.z));
            }
            for (int k = d.custom_gstart[term.y]; k < d.custom_gstart[term.y + 1]; ++k) {
                const double w = -d.custom_gweights[k] * fr;
                add_force(fa, N, d.custom_gatoms[k], make_double3(w * dv.x, w * dv.y, w * dv.z));
            }
        }
    }
}
