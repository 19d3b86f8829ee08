// Neighbour-list construction (K1) and the tiled direct-space pair kernel (K2).
#pragma once
#include "engine.cuh"

// ---------------------------------------------------------------------------------------------------------
// k_begin_eval: clear accumulators before a force evaluation and latch the rebuild request.
//   cm_mode: 0 keep, 1 zero cm_acc[parity], 2 zero cm_acc[parity] and flip parity (single-kernel steps)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_begin_eval(Dev d, int advance_noise, int advance_md, int cm_mode, int* cm_parity) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long nf = (long long)d.R * 3 * d.N;
    for (long long i = tid; i < nf; i += nthreads) d.f_env[i] = 0;
    if (d.n_alch > 0)
        for (long long i = tid; i < nf * ALCH_SLOTS; i += nthreads) d.f_alch[i] = 0;
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < d.R * N_ETERMS; i += blockDim.x) d.eacc[i] = 0;
        for (int i = threadIdx.x; i < d.R * ALCH_SLOTS * 3; i += blockDim.x) d.alch_acc[i] = 0;
        for (int i = threadIdx.x; i < d.R; i += blockDim.x) {
            Globals& g = d.g[i];
            g.do_rebuild = g.rebuild_request;
            g.rebuild_request = 0;
            g.noise_counter += advance_noise;
            g.md_counter += advance_md;
        }
        if (cm_mode) {
            const int p = *cm_parity;
            for (int i = threadIdx.x; i < d.R * 3; i += blockDim.x) d.cm_acc[(size_t)p * d.R * 3 + i] = 0;
            __syncthreads();
            if (cm_mode == 2 && threadIdx.x == 0) *cm_parity = p ^ 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_sort_atoms: one CTA per walker.  Counting sort of atoms by Morton-ranked cell, then builds the sorted
// mirrors, the inverse permutation and per-block bounding boxes.  Early exit unless a rebuild was latched.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wrap01(float f) { f -= floorf(f); return f >= 1.0f ? 0.0f : f; }

__global__ void __launch_bounds__(1024) k_sort_atoms(Dev d) {
    const int r = blockIdx.x;
    Globals& g = d.g[r];
    if (!g.do_rebuild) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = d.N, Npad = d.Npad, ncells = d.ncells;
    int* count = d.cell_count + (size_t)r * (ncells + 1);
    int* acell = d.atom_cell + (size_t)r * N;
    int* aslot = d.atom_slot + (size_t)r * N;
    const float4* posq = d.posq + (size_t)r * N;
    __shared__ int s_part[1024];
    __shared__ int s_total;

    for (int c = tid; c <= ncells; c += nt) count[c] = 0;
    __syncthreads();
    const float ibx = d.periodic ? d.boxf[3] : 0.f, iby = d.periodic ? d.boxf[4] : 0.f, ibz = d.periodic ? d.boxf[5] : 0.f;
    for (int a = tid; a < N; a += nt) {
        float4 p = posq[a];
        int cell = 0;
        if (d.periodic) {
            int cx = min((int)(wrap01(p.x * ibx) * d.ncell[0]), d.ncell[0] - 1);
            int cy = min((int)(wrap01(p.y * iby) * d.ncell[1]), d.ncell[1] - 1);
            int cz = min((int)(wrap01(p.z * ibz) * d.ncell[2]), d.ncell[2] - 1);
            cell = d.cell_order[(cx * d.ncell[1] + cy) * d.ncell[2] + cz];
        }
        acell[a] = cell;
        aslot[a] = atomicAdd(&count[cell], 1);
    }
    __syncthreads();
    // exclusive scan of count[0..ncells)
    const int per = (ncells + nt - 1) / nt;
    const int c0 = min(tid * per, ncells), c1 = min(c0 + per, ncells);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += count[c];
    s_part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < nt; off <<= 1) {
        int v = (tid >= off) ? s_part[tid - off] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int run = s_part[tid] - sum;
    for (int c = c0; c < c1; ++c) { int v = count[c]; count[c] = run; run += v; }
    if (tid == nt - 1) s_total = s_part[tid];
    __syncthreads();
    // scatter
    int* rank = d.rank + (size_t)r * N;
    float4* posq_s = d.posq_s + (size_t)r * Npad;
    float2* sigeps_s = d.sigeps_s + (size_t)r * Npad;
    int* orig_s = d.orig_s + (size_t)r * Npad;
    float4* pos_ref = d.pos_ref + (size_t)r * N;
    for (int a = tid; a < N; a += nt) {
        int s = count[acell[a]] + aslot[a];
        rank[a] = s;
        float4 p = posq[a];
        posq_s[s] = p;
        sigeps_s[s] = d.sigeps[a];
        orig_s[s] = a;
        pos_ref[a] = p;
    }
    const float qnan = __int_as_float(0x7fc00000);
    for (int s = N + tid; s < Npad; s += nt) {
        posq_s[s] = make_float4(qnan, qnan, qnan, 0.f);
        sigeps_s[s] = make_float2(0.f, 0.f);
        orig_s[s] = -1;
    }
    __syncthreads();
    // bounding boxes: one warp per block of 32 sorted atoms
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2];
    for (int b = warp; b < d.nblocks; b += nwarps) {
        float4 p = posq_s[b * 32 + lane];
        float4 p0;
        p0.x = __shfl_sync(0xffffffffu, p.x, 0);
        p0.y = __shfl_sync(0xffffffffu, p.y, 0);
        p0.z = __shfl_sync(0xffffffffu, p.z, 0);
        bool valid = (b * 32 + lane) < N;
        float dx = p.x - p0.x, dy = p.y - p0.y, dz = p.z - p0.z;
        if (d.periodic) {
            dx -= bx * rintf(dx * ibx);
            dy -= by * rintf(dy * iby);
            dz -= bz * rintf(dz * ibz);
        }
        const float big = 1e30f;
        float lox = warp_min(valid ? dx : big), hix = warp_max(valid ? dx : -big);
        float loy = warp_min(valid ? dy : big), hiy = warp_max(valid ? dy : -big);
        float loz = warp_min(valid ? dz : big), hiz = warp_max(valid ? dz : -big);
        if (lane == 0) {
            d.blk_center[(size_t)r * d.nblocks + b] = make_float4(p0.x + 0.5f * (lox + hix), p0.y + 0.5f * (loy + hiy),
                                                                   p0.z + 0.5f * (loz + hiz), 0.f);
            d.blk_half[(size_t)r * d.nblocks + b] = make_float4(0.5f * (hix - lox), 0.5f * (hiy - loy),
                                                                 0.5f * (hiz - loz), 0.f);
        }
    }
    if (tid == 0) {
        g.n_items = 0;
        g.item_overflow = 0;
        g.n_rebuilds += 1;
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_find_tiles: one warp per i-block.  Scans blocks j >= i, keeps j-atoms within the list cutoff of the
// i-block bounding box (warp-ballot compaction) and emits work items of up to TILE_CHUNKS*32 j-atoms together
// with per-lane exclusion bit masks.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_excluded(const Dev& d, int oi, ull wi, bool fari, int oj, bool farj) {
    if (oi < 0 || oj < 0) return false;
    int dd = oj - oi + 32;
    if ((unsigned)dd < 64u) return (wi >> dd) & 1ull;
    if (fari && farj) {
        long long code = oi < oj ? (long long)oi * d.N + oj : (long long)oj * d.N + oi;
        int lo = 0, hi = d.n_far - 1;
        while (lo <= hi) {
            int mid = (lo + hi) >> 1;
            long long v = d.far_codes[mid];
            if (v == code) return true;
            if (v < code) lo = mid + 1; else hi = mid - 1;
        }
    }
    return false;
}

#define FT_WARPS 4
__global__ void __launch_bounds__(FT_WARPS * 32) k_find_tiles(Dev d) {
    const int r = blockIdx.y;
    Globals& g = d.g[r];
    if (!g.do_rebuild) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bi = blockIdx.x * FT_WARPS + warp;
    if (bi >= d.nblocks) return;
    __shared__ int s_buf[FT_WARPS][ITEM_ATOMS + 32];
    int* buf = s_buf[warp];
    const int Npad = d.Npad, nb = d.nblocks;
    const float4* posq_s = d.posq_s + (size_t)r * Npad;
    const int* orig_s = d.orig_s + (size_t)r * Npad;
    const float4* bc = d.blk_center + (size_t)r * nb;
    const float4* bh = d.blk_half + (size_t)r * nb;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2];
    const float ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float4 ci = bc[bi], hi = bh[bi];
    const float cut2 = d.list_cutoff2;
    // i-lane data for exclusion masks
    const int oi = orig_s[bi * 32 + lane];
    const ull wi = oi >= 0 ? d.excl_win[oi] : 0ull;
    const bool fari = oi >= 0 ? d.has_far[oi] : false;
    int nbuf = 0;
    bool first_item = true;

    auto flush = [&](int count) {
        // emit one item holding `count` (<= ITEM_ATOMS) buffered atoms
        int item = 0;
        if (lane == 0) item = atomicAdd(&g.n_items, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= d.item_capacity) {
            if (lane == 0) g.item_overflow = 1;
            return;
        }
        const size_t ib = (size_t)r * d.item_capacity + item;
        int flags = first_item ? 256 : 0;
        for (int c = 0; c < TILE_CHUNKS; ++c) {
            int k = c * 32 + lane;
            int js = (k < count) ? buf[k] : -1;
            d.item_atoms[ib * ITEM_ATOMS + k] = js;
            int oj = js >= 0 ? orig_s[js] : -1;
            bool farj = oj >= 0 ? d.has_far[oj] : false;
            unsigned int mask = 0;
            if (c * 32 < count) {
                for (int l = 0; l < 32; ++l) {
                    int ojl = __shfl_sync(0xffffffffu, oj, l);
                    bool fjl = __shfl_sync(0xffffffffu, (int)farj, l);
                    if (pair_excluded(d, oi, wi, fari, ojl, fjl)) mask |= (1u << l);
                }
            }
            d.item_excl[ib * ITEM_ATOMS + k] = mask;
            if (__any_sync(0xffffffffu, mask != 0)) flags |= (1 << c);
        }
        if (lane == 0) {
            d.item_block[ib] = bi;
            d.item_natoms[ib] = count;
            d.item_flags[ib] = flags;
        }
        first_item = false;
    };

    // chunk 0 of the first item is the block's own atoms, in lane order
    buf[lane] = bi * 32 + lane;
    nbuf = 32;
    __syncwarp();
    for (int base = bi + 1; base < nb; base += 32) {
        int bj = base + lane;
        bool cand = false;
        if (bj < nb) {
            float4 cj = bc[bj], hj = bh[bj];
            float dx = ci.x - cj.x, dy = ci.y - cj.y, dz = ci.z - cj.z;
            if (d.periodic) {
                dx -= bx * rintf(dx * ibx);
                dy -= by * rintf(dy * iby);
                dz -= bz * rintf(dz * ibz);
            }
            dx = fmaxf(0.f, fabsf(dx) - hi.x - hj.x);
            dy = fmaxf(0.f, fabsf(dy) - hi.y - hj.y);
            dz = fmaxf(0.f, fabsf(dz) - hi.z - hj.z);
            cand = (dx * dx + dy * dy + dz * dz) < cut2;
        }
        unsigned int cmask = __ballot_sync(0xffffffffu, cand);
        while (cmask) {
            int l = __ffs(cmask) - 1;
            cmask &= cmask - 1;
            int bjj = base + l;
            int js = bjj * 32 + lane;
            float4 p = posq_s[js];
            float dx = p.x - ci.x, dy = p.y - ci.y, dz = p.z - ci.z;
            if (d.periodic) {
                dx -= bx * rintf(dx * ibx);
                dy -= by * rintf(dy * iby);
                dz -= bz * rintf(dz * ibz);
            }
            dx = fmaxf(0.f, fabsf(dx) - hi.x);
            dy = fmaxf(0.f, fabsf(dy) - hi.y);
            dz = fmaxf(0.f, fabsf(dz) - hi.z);
            bool in = (dx * dx + dy * dy + dz * dz) < cut2;   // NaN pads compare false
            unsigned int m = __ballot_sync(0xffffffffu, in);
            if (in) buf[nbuf + __popc(m & ((1u << lane) - 1u))] = js;
            nbuf += __popc(m);
            __syncwarp();
            if (nbuf >= ITEM_ATOMS) {
                flush(ITEM_ATOMS);
                int rem = nbuf - ITEM_ATOMS;
                int v = (lane < rem) ? buf[ITEM_ATOMS + lane] : 0;
                __syncwarp();
                if (lane < rem) buf[lane] = v;
                nbuf = rem;
                __syncwarp();
            }
        }
    }
    if (nbuf > 0) flush(nbuf);
}

// ---------------------------------------------------------------------------------------------------------
// k_pair: direct-space Lennard-Jones + Coulomb (Ewald erfc / reaction field / plain) on 32x32 tiles.
// One warp per work item.  Lane l owns i-atom l of the block (register resident); the j-atom package
// (x, y, z, q, sigma/2, 2 sqrt(eps), fx, fy, fz) rotates around the warp by shuffle so every lane meets
// every j once; i- and j-forces leave the warp as 64-bit fixed-point atomics (order independent → deterministic).
// ---------------------------------------------------------------------------------------------------------
#define NB_NOCUT 0
#define NB_RF 2
#define NB_PME 4

__device__ __forceinline__ float erfc_times(float ar, float expar) {
    // Abramowitz & Stegun 7.1.26: erfc(x) = poly(t) exp(-x^2), |error| <= 1.5e-7
    float t = __frcp_rn(1.0f + 0.3275911f * ar);
    return (0.254829592f + (-0.284496736f + (1.421413741f + (-1.453152027f + 1.061405429f * t) * t) * t) * t) * t * expar;
}

template <int METHOD, bool ENERGY>
__global__ void __launch_bounds__(256) k_pair(Dev d) {
    const int r = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int n_items = min(d.g[r].n_items, d.item_capacity);
    const int N = d.N, Npad = d.Npad;
    const float4* __restrict__ posq_s = d.posq_s + (size_t)r * Npad;
    const float2* __restrict__ sigeps_s = d.sigeps_s + (size_t)r * Npad;
    const int* __restrict__ orig_s = d.orig_s + (size_t)r * Npad;
    long long* fenv = d.f_env + (size_t)r * 3 * N;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2];
    const float ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float cut2 = METHOD == NB_NOCUT ? 3.0e38f : d.cutoff2;
    const float alpha = d.alpha, krf = d.krf, crf = d.crf;
    const float KE = (float)ONE_4PI_EPS0;
    float etot = 0.f;

    for (int item = blockIdx.x * wpb + (threadIdx.x >> 5); item < n_items; item += gridDim.x * wpb) {
        const size_t ib = (size_t)r * d.item_capacity + item;
        const int bi = d.item_block[ib];
        const int flags = d.item_flags[ib];
        const int count = d.item_natoms[ib];
        const float4 pi = posq_s[bi * 32 + lane];
        const float2 se_i = sigeps_s[bi * 32 + lane];
        const float qi = pi.w * KE;
        float fix = 0.f, fiy = 0.f, fiz = 0.f;
        for (int c = 0; c * 32 < count; ++c) {
            const int js = d.item_atoms[ib * ITEM_ATOMS + c * 32 + lane];
            const unsigned int excl = (flags >> c) & 1 ? d.item_excl[ib * ITEM_ATOMS + c * 32 + lane] : 0u;
            const bool self = (c == 0) && (flags & 256);
            float4 pj;
            float2 se_j;
            if (js >= 0) { pj = posq_s[js]; se_j = sigeps_s[js]; }
            else { pj = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f); se_j = make_float2(0.f, 0.f); }
            float fjx = 0.f, fjy = 0.f, fjz = 0.f;
            // before iteration k the package in this lane originates from lane (lane + k) & 31
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const int src = (lane + k) & 31;
                float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                if (METHOD != NB_NOCUT) {
                    dx -= bx * rintf(dx * ibx);
                    dy -= by * rintf(dy * iby);
                    dz -= bz * rintf(dz * ibz);
                }
                const float r2 = dx * dx + dy * dy + dz * dz;
                bool ok = r2 < cut2;
                ok = ok && !((excl >> src) & 1u);
                if (self) ok = ok && (src > lane);
                if (ok) {
                    const float invr = rsqrtf(r2);
                    const float invr2 = invr * invr;
                    const float sig = se_i.x + se_j.x;
                    const float s2 = sig * sig * invr2;
                    const float s6 = s2 * s2 * s2;
                    const float eps4 = se_i.y * se_j.y;
                    float de = eps4 * (12.0f * s6 * s6 - 6.0f * s6);
                    const float qq = qi * pj.w;
                    if (METHOD == NB_PME) {
                        const float rr = r2 * invr;
                        const float ar = alpha * rr;
                        const float ex = expf(-ar * ar);
                        const float ec = erfc_times(ar, ex);
                        de += qq * invr * (ec + (float)TWO_OVER_SQRT_PI * ar * ex);
                        if (ENERGY) etot += eps4 * (s6 * s6 - s6) + qq * invr * ec;
                    } else if (METHOD == NB_RF) {
                        de += qq * (invr - 2.0f * krf * r2);
                        if (ENERGY) etot += eps4 * (s6 * s6 - s6) + qq * (invr + krf * r2 - crf);
                    } else {
                        de += qq * invr;
                        if (ENERGY) etot += eps4 * (s6 * s6 - s6) + qq * invr;
                    }
                    de *= invr2;
                    dx *= de; dy *= de; dz *= de;
                    fix += dx; fiy += dy; fiz += dz;
                    fjx -= dx; fjy -= dy; fjz -= dz;
                }
                // rotate the j package to the previous lane (so this lane next sees lane+k+1's atom)
                const int from = (lane + 1) & 31;
                pj.x = __shfl_sync(0xffffffffu, pj.x, from);
                pj.y = __shfl_sync(0xffffffffu, pj.y, from);
                pj.z = __shfl_sync(0xffffffffu, pj.z, from);
                pj.w = __shfl_sync(0xffffffffu, pj.w, from);
                se_j.x = __shfl_sync(0xffffffffu, se_j.x, from);
                se_j.y = __shfl_sync(0xffffffffu, se_j.y, from);
                fjx = __shfl_sync(0xffffffffu, fjx, from);
                fjy = __shfl_sync(0xffffffffu, fjy, from);
                fjz = __shfl_sync(0xffffffffu, fjz, from);
            }
            // after 32 rotations the package is back in its owner lane
            if (js >= 0 && (fjx != 0.f || fjy != 0.f || fjz != 0.f)) {
                const int oj = orig_s[js];
                fx_addf(&fenv[oj], fjx, (float)FORCE_SCALE);
                fx_addf(&fenv[N + oj], fjy, (float)FORCE_SCALE);
                fx_addf(&fenv[2 * N + oj], fjz, (float)FORCE_SCALE);
            }
        }
        const int oi = orig_s[bi * 32 + lane];
        if (oi >= 0) {
            fx_addf(&fenv[oi], fix, (float)FORCE_SCALE);
            fx_addf(&fenv[N + oi], fiy, (float)FORCE_SCALE);
            fx_addf(&fenv[2 * N + oi], fiz, (float)FORCE_SCALE);
        }
    }
    if (ENERGY) {
        float e = warp_sum(etot);
        if (lane == 0 && e != 0.f) fx_add(&d.eacc[r * N_ETERMS + E_PAIR], (double)e, ENERGY_SCALE);
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_neighbor_pairs: enumerate (for tests) the non-excluded pairs within the cutoff found through the tile list.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_neighbor_pairs(Dev d, int r, long long* codes, unsigned long long capacity, unsigned long long* n_out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int n_items = min(d.g[r].n_items, d.item_capacity);
    const int Npad = d.Npad;
    const float4* posq_s = d.posq_s + (size_t)r * Npad;
    const int* orig_s = d.orig_s + (size_t)r * Npad;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2];
    const float ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float cut2 = d.periodic ? d.cutoff2 : 3.0e38f;
    for (int item = blockIdx.x * wpb + (threadIdx.x >> 5); item < n_items; item += gridDim.x * wpb) {
        const size_t ib = (size_t)r * d.item_capacity + item;
        const int bi = d.item_block[ib];
        const int flags = d.item_flags[ib];
        const int count = d.item_natoms[ib];
        const float4 pi = posq_s[bi * 32 + lane];
        const int oi = orig_s[bi * 32 + lane];
        for (int c = 0; c * 32 < count; ++c) {
            const int js = d.item_atoms[ib * ITEM_ATOMS + c * 32 + lane];
            const unsigned int excl = d.item_excl[ib * ITEM_ATOMS + c * 32 + lane];
            const bool self = (c == 0) && (flags & 256);
            for (int l = 0; l < 32; ++l) {
                int jsl = __shfl_sync(0xffffffffu, js, l);
                if (jsl < 0) continue;
                float4 pj = posq_s[jsl];
                int oj = orig_s[jsl];
                float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                if (d.periodic) {
                    dx -= bx * rintf(dx * ibx);
                    dy -= by * rintf(dy * iby);
                    dz -= bz * rintf(dz * ibz);
                }
                bool ok = (dx * dx + dy * dy + dz * dz) < cut2 && oi >= 0 && !((excl >> l) & 1u);
                if (self) ok = ok && (l > lane);
                if (ok) {
                    unsigned long long slot = atomicAdd(n_out, 1ull);
                    if (slot < capacity) codes[slot] = oi < oj ? (long long)oi * d.N + oj : (long long)oj * d.N + oi;
                }
            }
        }
    }
}
