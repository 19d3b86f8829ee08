// Neighbour-list construction (K1) and the tiled direct-space pair kernel (K2).
#pragma once
#include <cooperative_groups.h>
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
            // an inner-list refresh is due when some atom moved > inner skin / 2 since the last prune; it becomes a
            // full rebuild when, at that moment, some atom sits > (outer - inner skin) / 2 away from the outer reference
            g.do_rebuild = g.rebuild_request == 2 || (g.prune_request && g.rebuild_request);
            g.do_prune = g.prune_request || g.do_rebuild;
            g.rebuild_request = 0;
            g.prune_request = 0;
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
// k_sort_atoms: one CTA per walker.  Deterministic counting sort of the atoms by Morton-ranked cell (cell edge >=
// half the list cutoff), atoms inside a cell ordered by topology index so that every later accumulation order —
// and therefore every float sum — is reproducible.  Builds the sorted mirrors and the inverse permutation.
// Early exit unless a rebuild was latched by k_begin_eval.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wrap01(float f) { f -= floorf(f); return f >= 1.0f ? 0.0f : f; }

__device__ __forceinline__ void atom_cell_coords(const Dev& d, float4 p, int& cx, int& cy, int& cz) {
    // non-finite coordinates (a walker that blew up) must not index out of bounds: clamp into the grid
    cx = max(0, min((int)(wrap01(p.x * d.boxf[3]) * d.ncell[0]), d.ncell[0] - 1));
    cy = max(0, min((int)(wrap01(p.y * d.boxf[4]) * d.ncell[1]), d.ncell[1] - 1));
    cz = max(0, min((int)(wrap01(p.z * d.boxf[5]) * d.ncell[2]), d.ncell[2] - 1));
}

#define SORT_CTAS 8
__global__ void __cluster_dims__(SORT_CTAS, 1, 1) __launch_bounds__(1024) k_sort_atoms(Dev d) {
    // one thread-block cluster (8 CTAs, hardware cluster barrier) per walker
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int r = blockIdx.y;
    Globals& g = d.g[r];
    if (!g.do_rebuild) return;                 // uniform over the cluster
    const int cta = (int)cluster.block_rank();
    const int tid = cta * blockDim.x + threadIdx.x, nt = SORT_CTAS * blockDim.x;
    const int lane = threadIdx.x & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int N = d.N, Npad = d.Npad, ncells = d.ncells;
    int* start = d.cell_start + (size_t)r * (ncells + 1);
    int* cursor = d.cell_cursor + (size_t)r * (ncells + 1);
    int* acell = d.atom_cell + (size_t)r * N;
    const float4* posq = d.posq + (size_t)r * N;
    int* orig_s = d.orig_s + (size_t)r * Npad;
    __shared__ int s_part[1024];

    for (int c = tid; c <= ncells; c += nt) cursor[c] = 0;
    cluster.sync();
    for (int a = tid; a < N; a += nt) {
        int cell = 0;
        if (d.periodic) {
            int cx, cy, cz;
            atom_cell_coords(d, posq[a], cx, cy, cz);
            cell = (cx * d.ncell[1] + cy) * d.ncell[2] + cz;       // row-major: a z-column of cells is contiguous
        }
        acell[a] = cell;
        atomicAdd(&cursor[cell], 1);
    }
    cluster.sync();
    if (cta == 0) {
        // exclusive scan of the per-cell counts by the first CTA
        const int t = threadIdx.x, n1 = blockDim.x;
        const int per = (ncells + n1 - 1) / n1;
        const int c0 = min(t * per, ncells), c1 = min(c0 + per, ncells);
        int sum = 0;
        for (int c = c0; c < c1; ++c) sum += cursor[c];
        s_part[t] = sum;
        __syncthreads();
        for (int off = 1; off < n1; off <<= 1) {
            int v = (t >= off) ? s_part[t - off] : 0;
            __syncthreads();
            s_part[t] += v;
            __syncthreads();
        }
        int run = s_part[t] - sum;
        for (int c = c0; c < c1; ++c) { int v = cursor[c]; start[c] = run; cursor[c] = run; run += v; }
        if (t == n1 - 1) start[ncells] = N;
    }
    cluster.sync();
    // scatter (arbitrary order inside a cell) ...
    for (int a = tid; a < N; a += nt) orig_s[atomicAdd(&cursor[acell[a]], 1)] = a;
    cluster.sync();
    // ... then order every cell's segment by topology index: one warp per cell, rank by counting smaller keys
    for (int c = warp; c < ncells; c += nwarps) {
        const int s0 = start[c], n = start[c + 1] - s0;
        if (n <= 1) continue;
        if (n <= 64) {
            const int v0 = lane < n ? orig_s[s0 + lane] : 0x7fffffff;
            const int v1 = lane + 32 < n ? orig_s[s0 + 32 + lane] : 0x7fffffff;
            int r0 = 0, r1 = 0;
            for (int m = 0; m < 32; ++m) {
                const int a0 = __shfl_sync(0xffffffffu, v0, m), a1 = __shfl_sync(0xffffffffu, v1, m);
                r0 += (a0 < v0) + (a1 < v0);
                r1 += (a0 < v1) + (a1 < v1);
            }
            __syncwarp();
            if (lane < n) orig_s[s0 + r0] = v0;
            if (lane + 32 < n) orig_s[s0 + r1] = v1;
        } else if (lane == 0) {
            for (int i = s0 + 1; i < s0 + n; ++i) {
                const int v = orig_s[i];
                int j = i - 1;
                while (j >= s0 && orig_s[j] > v) { orig_s[j + 1] = orig_s[j]; --j; }
                orig_s[j + 1] = v;
            }
        }
    }
    cluster.sync();
    int* rank = d.rank + (size_t)r * N;
    float4* posq_s = d.posq_s + (size_t)r * Npad;
    float2* sigeps_s = d.sigeps_s + (size_t)r * Npad;
    float4* pos_ref = d.pos_ref_outer + (size_t)r * N;
    for (int s = tid; s < N; s += nt) {
        const int a = orig_s[s];
        rank[a] = s;
        const float4 p = posq[a];
        posq_s[s] = p;
        sigeps_s[s] = d.sigeps[a];
        pos_ref[a] = p;
    }
    const float qnan = __int_as_float(0x7fc00000);
    for (int s = N + tid; s < Npad; s += nt) {
        posq_s[s] = make_float4(qnan, qnan, qnan, 0.f);
        sigeps_s[s] = make_float2(0.f, 0.f);
        orig_s[s] = -1;
    }
    if (tid == 0) {
        g.item_overflow = 0;
        g.n_rebuilds += 1;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Exclusion test on topology indices: a 64-bit window mask covers partners within +-32, a sorted code list the rest.
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

// ---------------------------------------------------------------------------------------------------------
// k_build_list: outer Verlet list (cutoff + outer skin) by cell search.  Full list (i sees j and j sees i): the pair
// kernel needs no j-side force scatter and no exclusion test.
//
// One warp owns BUILD_GROUP = 8 consecutive atoms of the cell-sorted order (a compact group: one cell, sometimes two);
// lane = (atom a = lane & 7, candidate subset q = lane >> 3).  The warp walks the cell columns around the group's
// bounding box; candidates are staged 32 at a time in shared memory, already shifted to the periodic image that can
// be in range, and in iteration t lane (a, q) tests candidate 8 q + t against atom a: one LDS and ~12 ALU instructions
// per 32 distance tests, no per-candidate global load.  Exclusions are only looked at in chunks that hold a candidate
// whose topology index is near the group's (warp-uniform test), i.e. almost never.  Survivors go to a per-lane ring in
// shared memory (bank-skewed) that the whole warp flushes 32 entries at a time, so global stores stay coalesced.
// An atom's outer list is therefore stored as BUILD_SUB = 4 sub-rows (one per candidate subset); scan order is fixed,
// hence so is the order of every sub-row (reproducible float sums downstream).
// ---------------------------------------------------------------------------------------------------------
#define NL_LANES 8
#define NL_BLOCK 128
#define BUILD_WARPS 4
#define BUILD_RING 64
#define BUILD_GROUP 8
#define BUILD_SUB 4

__device__ __forceinline__ void sts_u32(unsigned int addr, unsigned int v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <typename IDX>
__device__ __forceinline__ void build_flush(IDX* __restrict__ rows, int Mq, const unsigned int* ring, int lane,
                                            int cnt, int& flushed) {
    // rows: sub-row 0 of the group's first atom; lane L = (atom L & 7, subset L >> 3) owns sub-row (L & 7) * 4 + (L >> 3)
    unsigned int full = __ballot_sync(0xffffffffu, cnt - flushed >= 32);
    while (full) {
        const int L = __ffs(full) - 1;
        full &= full - 1;
        const int cL = __shfl_sync(0xffffffffu, flushed, L);
        const unsigned int v = ring[L * BUILD_RING + ((cL + lane + L) & (BUILD_RING - 1))];
        if (cL < Mq) rows[(size_t)((L & (BUILD_GROUP - 1)) * BUILD_SUB + (L >> 3)) * Mq + cL + lane] = (IDX)v;   // Mq % 32 == 0
        if (lane == L) flushed += 32;
    }
}

template <bool RINT, typename IDX>
__device__ __forceinline__ void build_scan_run(const Dev& d, const float4* __restrict__ posq_s, const int* __restrict__ orig_s,
                                               int s0, int s1, float sx, float sy, float sz, bool rx, bool ry, bool rz,
                                               float4* cand, unsigned int* ring, IDX* rows, int lane, float4 pi, int oi,
                                               ull wi, bool fari, bool anyfar, const int (&og)[BUILD_GROUP],
                                               const unsigned int (&osp)[BUILD_GROUP], int& cnt, int& flushed) {
    const float cut2 = d.outer_cutoff2;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2], ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float qnan = __int_as_float(0x7fc00000);
    const int q = lane >> 3;
    unsigned int* myring = ring + lane * BUILD_RING;
    const unsigned int ring_addr = (unsigned int)__cvta_generic_to_shared(myring), lane4 = 4u * lane;
    for (int base = s0; base < s1; base += 32) {
        bool near = false;
        {
            const int s = base + lane;
            float4 c = make_float4(qnan, qnan, qnan, 0.f);
            if (s < s1) {
                c = posq_s[s];
                c.x += sx; c.y += sy; c.z += sz;
                const int oj = orig_s[s];
                c.w = __int_as_float(oj);
#pragma unroll
                for (int k = 0; k < BUILD_GROUP; ++k)               // inside the exclusion window of a group atom
                    near = near || (unsigned int)(oj - og[k]) <= osp[k];
            }
            __syncwarp();
            cand[lane + (lane >> 3)] = c;                           // 8-candidate pieces, padded: conflict-free LDS
            __syncwarp();
        }
        const bool check = anyfar || __any_sync(0xffffffffu, near);   // warp-uniform
        if (!check) {
            // fast path: no candidate of this chunk can be excluded from (or be) an atom of the group.  All eight
            // candidates are loaded and measured before the first ring store, so the eight chains overlap.
            float r2v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float4 c = cand[q * 9 + t];                  // candidate 8 q + t
                float dx = c.x - pi.x, dy = c.y - pi.y, dz = c.z - pi.z;
                if (RINT) {
                    if (rx) dx -= bx * rintf(dx * ibx);
                    if (ry) dy -= by * rintf(dy * iby);
                    if (rz) dz -= bz * rintf(dz * ibz);
                }
                r2v[t] = dx * dx + dy * dy + dz * dz;
            }
            const unsigned int vq = (unsigned int)(base + 8 * q);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (r2v[t] < cut2) {                               // NaN (padding) compares false
                    // ring slot (cnt + lane) & 63 of this lane's 256-byte aligned ring
                    sts_u32(ring_addr | ((((unsigned int)cnt << 2) + lane4) & 0xFFu), vq + t);
                    ++cnt;
                }
            }
        } else {
#pragma unroll 1
            for (int t = 0; t < 8; ++t) {
                const float4 c = cand[q * 9 + t];
                float dx = c.x - pi.x, dy = c.y - pi.y, dz = c.z - pi.z;
                if (RINT) {
                    if (rx) dx -= bx * rintf(dx * ibx);
                    if (ry) dy -= by * rintf(dy * iby);
                    if (rz) dz -= bz * rintf(dz * ibz);
                }
                if (dx * dx + dy * dy + dz * dz < cut2) {
                    const int oj = __float_as_int(c.w);
                    const unsigned int dd = (unsigned int)(oj - oi + 32);
                    bool ok = true;
                    if (dd < 64u) ok = !((wi >> dd) & 1ull);           // includes the atom itself (bit 32)
                    else if (fari) ok = !pair_excluded(d, oi, wi, true, oj, d.has_far[oj]);   // rare
                    if (ok) {
                        myring[(cnt + lane) & (BUILD_RING - 1)] = (unsigned int)(base + 8 * q + t);
                        ++cnt;
                    }
                }
            }
        }
        __syncwarp();
        build_flush<IDX>(rows, d.nlo_M, ring, lane, cnt, flushed);
    }
}

template <bool RINT, typename IDX>
__device__ __forceinline__ void build_scan_cells(const Dev& d, const float4* __restrict__ posq_s, const int* __restrict__ orig_s,
                                              const int* __restrict__ start, bool rx, bool ry, bool rz, int x0, int x1,
                                              int y0, int y1, int za, int zb, float lox, float hix, float loy, float hiy,
                                              float loz, float hiz, float4* cand, unsigned int* ring, IDX* rows, int lane,
                                              float4 pi, int oi, ull wi, bool fari, bool anyfar,
                                              const int (&og)[BUILD_GROUP], const unsigned int (&osp)[BUILD_GROUP], int& cnt,
                                              int& flushed) {
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2];
    const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
    const float ex = bx / ncx, ey = by / ncy, ez = bz / ncz;
    const float cut2 = d.outer_cutoff2;
    for (int rxc = x0; rxc <= x1; ++rxc) {
        int ax = rxc;
        float sx = 0.f, dxc = 0.f;
        if (!rx) {
            if (ax < 0) { ax += ncx; sx = -bx; } else if (ax >= ncx) { ax -= ncx; sx = bx; }
            dxc = fmaxf(0.f, fmaxf(rxc * ex - hix, lox - (rxc + 1) * ex));   // gap between the group and the slab
        }
        for (int ryc = y0; ryc <= y1; ++ryc) {
            int ay = ryc;
            float sy = 0.f, dyc = 0.f;
            if (!ry) {
                if (ay < 0) { ay += ncy; sy = -by; } else if (ay >= ncy) { ay -= ncy; sy = by; }
                dyc = fmaxf(0.f, fmaxf(ryc * ey - hiy, loy - (ryc + 1) * ey));
            }
            const float rem2 = cut2 - dxc * dxc - dyc * dyc;
            if (rem2 <= 0.f) continue;                                        // column out of reach
            int z0 = 0, z1 = ncz - 1;
            if (!rz) {
                const float zr = sqrtf(rem2);
                z0 = max(za - 2, (int)floorf((loz - zr) / ez));
                z1 = min(zb + 2, (int)floorf((hiz + zr) / ez));
            }
            const int row = (ax * ncy + ay) * ncz;
#pragma unroll 1
            for (int seg = 0; seg < 3; ++seg) {
                int a0, a1;
                float sz = 0.f;
                if (seg == 0) { a0 = max(z0, 0); a1 = min(z1, ncz - 1); }
                else if (seg == 1) { if (z0 >= 0) continue; a0 = z0 + ncz; a1 = ncz - 1; sz = -bz; }
                else { if (z1 < ncz) continue; a0 = 0; a1 = z1 - ncz; sz = bz; }
                if (a0 > a1) continue;
                const int s0 = start[row + a0], s1 = start[row + a1 + 1];
                build_scan_run<RINT, IDX>(d, posq_s, orig_s, s0, s1, sx, sy, sz, rx, ry, rz, cand, ring, rows, lane,
                                          pi, oi, wi, fari, anyfar, og, osp, cnt, flushed);
            }
        }
    }
}

template <typename IDX>
__global__ void __launch_bounds__(BUILD_WARPS * 32) k_build_list(Dev d) {
    const int r = blockIdx.y;
    Globals& g = d.g[r];
    if (!g.do_rebuild) return;
    __shared__ float4 s_cand[BUILD_WARPS][36];                   // 4 pieces of 8 candidates + 1 pad each
    __shared__ __align__(256) unsigned int s_ring[BUILD_WARPS][32 * BUILD_RING];    // per lane: 64 entries = 256 B
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i0 = (blockIdx.x * BUILD_WARPS + w) * BUILD_GROUP, i = i0 + (lane & (BUILD_GROUP - 1));
    const int N = d.N, Npad = d.Npad;
    if (i0 >= Npad) return;
    int* counts = d.nlo_count + ((size_t)r * Npad + i0) * BUILD_SUB;        // [atom of the group][subset]
    if (i0 >= N) { counts[(lane & (BUILD_GROUP - 1)) * BUILD_SUB + (lane >> 3)] = 0; return; }
    const float4* __restrict__ posq_s = d.posq_s + (size_t)r * Npad;
    const int* __restrict__ orig_s = d.orig_s + (size_t)r * Npad;
    const int* __restrict__ start = d.cell_start + (size_t)r * (d.ncells + 1);
    IDX* rows = reinterpret_cast<IDX*>(d.nlo_list) + ((size_t)r * Npad + i0) * BUILD_SUB * d.nlo_M;
    float4* cand = s_cand[w];
    unsigned int* ring = s_ring[w];
    const bool valid = i < N;
    const float4 pi = posq_s[i];                                 // padding rows hold NaN: never in range
    const int o0 = __shfl_sync(0xffffffffu, valid ? orig_s[i] : 0, 0);      // lane 0 is always a real atom
    const int oi = valid ? orig_s[i] : o0;
    const ull wi = valid ? (d.excl_win[oi] | (1ull << 32)) : 0ull;
    const bool fari = valid ? d.has_far[oi] : false;
    const bool anyfar = __any_sync(0xffffffffu, fari);
    // exclusion window of every atom of the group as [og, og + osp] in topology indices (waters: their own molecule)
    int og[BUILD_GROUP];
    unsigned int osp[BUILD_GROUP];
    {
        const int below = valid ? 32 - (__ffsll((long long)wi) - 1) : 0, above = valid ? 31 - __clzll((long long)wi) : 0;
#pragma unroll
        for (int k = 0; k < BUILD_GROUP; ++k) {
            og[k] = __shfl_sync(0xffffffffu, oi - below, k);
            osp[k] = (unsigned int)__shfl_sync(0xffffffffu, below + above, k);
        }
    }
    int cnt = 0, flushed = 0;
    if (!d.periodic) {
        build_scan_run<false, IDX>(d, posq_s, orig_s, 0, N, 0.f, 0.f, 0.f, false, false, false, cand, ring, rows, lane,
                                   pi, oi, wi, fari, anyfar, og, osp, cnt, flushed);
    } else {
        const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
        // bounding box of the group, in cells and in space (padding lanes copy lane 0)
        const float l0x = __shfl_sync(0xffffffffu, pi.x, 0), l0y = __shfl_sync(0xffffffffu, pi.y, 0);
        const float l0z = __shfl_sync(0xffffffffu, pi.z, 0);
        const float4 p0 = valid ? pi : make_float4(l0x, l0y, l0z, 0.f);
        int cx, cy, cz;
        atom_cell_coords(d, p0, cx, cy, cz);
        const int xa = __reduce_min_sync(0xffffffffu, cx), xb = __reduce_max_sync(0xffffffffu, cx);
        const int ya = __reduce_min_sync(0xffffffffu, cy), yb = __reduce_max_sync(0xffffffffu, cy);
        const int za = __reduce_min_sync(0xffffffffu, cz), zb = __reduce_max_sync(0xffffffffu, cz);
        const float lox = warp_min(p0.x), hix = warp_max(p0.x), loy = warp_min(p0.y), hiy = warp_max(p0.y);
        const float loz = warp_min(p0.z), hiz = warp_max(p0.z);
        // a dimension whose scan range would cover a cell twice is scanned once, with the rint() minimum image
        const bool rx = xb - xa + 5 > ncx, ry = yb - ya + 5 > ncy, rz = zb - za + 5 > ncz;
        const int x0 = rx ? 0 : xa - 2, x1 = rx ? ncx - 1 : xb + 2;
        const int y0 = ry ? 0 : ya - 2, y1 = ry ? ncy - 1 : yb + 2;
        if (rx || ry || rz)
            build_scan_cells<true, IDX>(d, posq_s, orig_s, start, rx, ry, rz, x0, x1, y0, y1, za, zb, lox, hix, loy, hiy,
                                        loz, hiz, cand, ring, rows, lane, pi, oi, wi, fari, anyfar, og, osp, cnt, flushed);
        else
            build_scan_cells<false, IDX>(d, posq_s, orig_s, start, false, false, false, x0, x1, y0, y1, za, zb, lox, hix,
                                         loy, hiy, loz, hiz, cand, ring, rows, lane, pi, oi, wi, fari, anyfar, og, osp, cnt, flushed);
    }
    // drain the rings: the warp writes each lane's remaining (< 32) entries
    __syncwarp();
    for (int L = 0; L < 32; ++L) {
        const int cL = __shfl_sync(0xffffffffu, flushed, L), nL = __shfl_sync(0xffffffffu, cnt, L) - cL;
        if (lane < nL && cL + lane < d.nlo_M)
            rows[(size_t)((L & (BUILD_GROUP - 1)) * BUILD_SUB + (L >> 3)) * d.nlo_M + cL + lane] =
                (IDX)ring[L * BUILD_RING + ((cL + lane + L) & (BUILD_RING - 1))];
    }
    if (cnt > d.nlo_M) { g.item_overflow = 1; cnt = d.nlo_M; }
    counts[(lane & (BUILD_GROUP - 1)) * BUILD_SUB + (lane >> 3)] = cnt;
}

// ---------------------------------------------------------------------------------------------------------
// k_prune_list: refresh the inner Verlet list (cutoff + inner skin) from the outer one (cutoff + outer skin) at the
// current coordinates: one warp per atom streams the outer list (coalesced), keeps the entries inside the inner list
// cutoff with ordered ballot compaction, and records the reference positions of the inner list.
// Runs every few steps; the expensive cell search (k_sort_atoms + k_build_list) only every ~10-20 steps.
// ---------------------------------------------------------------------------------------------------------
template <typename IDX>
__global__ void __launch_bounds__(128) k_prune_list(Dev d) {
    const int r = blockIdx.y;
    Globals& g = d.g[r];
    if (!g.do_prune) return;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int N = d.N, Npad = d.Npad;
    if (i >= Npad) return;
    const float4* __restrict__ posq_s = d.posq_s + (size_t)r * Npad;
    const IDX* __restrict__ outer_rows = reinterpret_cast<const IDX*>(d.nlo_list) + ((size_t)r * Npad + i) * BUILD_SUB * d.nlo_M;
    IDX* inner = reinterpret_cast<IDX*>(d.nl_list) + ((size_t)r * Npad + i) * d.nl_M;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2], ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float cut2 = d.list_cutoff2;
    const float4 pi = posq_s[i];
    int cnt = 0;
    for (int sub = 0; sub < BUILD_SUB; ++sub) {
        const IDX* __restrict__ outer = outer_rows + (size_t)sub * d.nlo_M;
        const int n_outer = i < N ? d.nlo_count[((size_t)r * Npad + i) * BUILD_SUB + sub] : 0;
        // two chunks of 32 entries per iteration: twice the gathers in flight (the kernel is latency bound)
        for (int base = 0; base < n_outer; base += 64) {
            const int k0 = base + lane, k1 = k0 + 32;
            bool ok0 = k0 < n_outer, ok1 = k1 < n_outer;
            const int s0 = ok0 ? (int)outer[k0] : i, s1 = ok1 ? (int)outer[k1] : i;
            const float4 pj0 = posq_s[s0], pj1 = posq_s[s1];
            float dx0 = pi.x - pj0.x, dy0 = pi.y - pj0.y, dz0 = pi.z - pj0.z;
            float dx1 = pi.x - pj1.x, dy1 = pi.y - pj1.y, dz1 = pi.z - pj1.z;
            if (d.periodic) {
                dx0 -= bx * rintf(dx0 * ibx); dy0 -= by * rintf(dy0 * iby); dz0 -= bz * rintf(dz0 * ibz);
                dx1 -= bx * rintf(dx1 * ibx); dy1 -= by * rintf(dy1 * iby); dz1 -= bz * rintf(dz1 * ibz);
            }
            ok0 = ok0 && (dx0 * dx0 + dy0 * dy0 + dz0 * dz0) < cut2;
            ok1 = ok1 && (dx1 * dx1 + dy1 * dy1 + dz1 * dz1) < cut2;
            const unsigned int m0 = __ballot_sync(0xffffffffu, ok0), m1 = __ballot_sync(0xffffffffu, ok1);
            const unsigned int below = (1u << lane) - 1u;
            if (ok0) {
                const int slot = cnt + __popc(m0 & below);
                if (slot < d.nl_M) inner[slot] = (IDX)s0;
            }
            cnt += __popc(m0);
            if (ok1) {
                const int slot = cnt + __popc(m1 & below);
                if (slot < d.nl_M) inner[slot] = (IDX)s1;
            }
            cnt += __popc(m1);
        }
    }
    if (cnt > d.nl_M) { g.item_overflow = 1; cnt = d.nl_M; }
    if (lane == 0) {
        d.nl_count[(size_t)r * Npad + i] = cnt;
        if (i < N) d.pos_ref[(size_t)r * N + d.orig_s[(size_t)r * Npad + i]] = pi;     // inner-list reference positions
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_pair: direct-space Lennard-Jones + Coulomb (Ewald erfc / reaction field / plain) over the Verlet list.
// NL_LANES lanes share one i-atom (register resident) and stride over its neighbour list with coalesced index
// loads; j data are gathered from the spatially sorted float4 mirror (L1/L2 resident); the partial forces are
// combined by warp shuffles in a fixed order and leave the warp as 64-bit fixed-point atomics.
// ---------------------------------------------------------------------------------------------------------
#define NB_NOCUT 0
#define NB_RF 2
#define NB_PME 4

__device__ __forceinline__ float erfc_times(float ar, float expar) {
    // Abramowitz & Stegun 7.1.26: erfc(x) = poly(t) exp(-x^2), |error| <= 1.5e-7
    const float t = __fdividef(1.0f, 1.0f + 0.3275911f * ar);
    return (0.254829592f + (-0.284496736f + (1.421413741f + (-1.453152027f + 1.061405429f * t) * t) * t) * t) * t * expar;
}

template <int METHOD, bool ENERGY, typename IDX>
__global__ void __launch_bounds__(NL_BLOCK) k_pair(Dev d) {
    const int r = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int part = lane & (NL_LANES - 1);
    const int i = (blockIdx.x * NL_BLOCK + threadIdx.x) / NL_LANES;
    const int N = d.N, Npad = d.Npad;
    if (i >= Npad) return;
    const float4* __restrict__ posq_s = d.posq_s + (size_t)r * Npad;
    const float2* __restrict__ sigeps_s = d.sigeps_s + (size_t)r * Npad;
    const IDX* __restrict__ list = reinterpret_cast<const IDX*>(d.nl_list) + ((size_t)r * Npad + i) * d.nl_M;
    const int cnt = d.nl_count[(size_t)r * Npad + i];
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2];
    const float ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float cut2 = METHOD == NB_NOCUT ? 3.0e38f : d.cutoff2;
    const float alpha = d.alpha, krf = d.krf, crf = d.crf;
    const float4 pi = posq_s[i];
    const float2 se_i = sigeps_s[i];
    const float qi = pi.w * (float)ONE_4PI_EPS0;
    float fx = 0.f, fy = 0.f, fz = 0.f, etot = 0.f;
#pragma unroll 4
    for (int k = part; k < cnt; k += NL_LANES) {
        const int s = (int)list[k];
        const float4 pj = posq_s[s];
        const float2 se_j = sigeps_s[s];
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        if (METHOD != NB_NOCUT) {
            dx -= bx * rintf(dx * ibx);
            dy -= by * rintf(dy * iby);
            dz -= bz * rintf(dz * ibz);
        }
        const float r2 = dx * dx + dy * dy + dz * dz;
        // branch-free: entries in the skin shell are computed and masked, so the loads of the unrolled iterations
        // are issued together instead of behind a divergent branch
        const float in = r2 < cut2 ? 1.0f : 0.0f;
        const float invr = rsqrtf(r2);
        const float invr2 = invr * invr;
        const float sig = se_i.x + se_j.x;
        const float s2 = sig * sig * invr2;
        const float s6 = s2 * s2 * s2;
        const float eps4 = se_i.y * se_j.y * in;
        float de = eps4 * (12.0f * s6 * s6 - 6.0f * s6);
        const float qq = qi * pj.w * in;
        if (METHOD == NB_PME) {
            const float ar = alpha * r2 * invr;
            const float ex = __expf(-ar * ar);
            const float ec = erfc_times(ar, ex);
            de += qq * invr * (ec + (float)TWO_OVER_SQRT_PI * ar * ex);
            if (ENERGY) etot += eps4 * (s6 * s6 - s6) + qq * invr * ec;
        } else if (METHOD == NB_RF) {
            de += qq * (invr - 2.0f * krf * r2);
            if (ENERGY) etot += eps4 * (s6 * s6 - s6) + qq * (invr + krf * r2 - crf) ;
        } else {
            de += qq * invr;
            if (ENERGY) etot += eps4 * (s6 * s6 - s6) + qq * invr;
        }
        de *= invr2;
        fx += dx * de; fy += dy * de; fz += dz * de;
    }
    // combine the NL_LANES partial sums (fixed xor tree → deterministic)
#pragma unroll
    for (int o = NL_LANES / 2; o > 0; o >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if (part == 0 && i < N) {
        const int oi = d.orig_s[(size_t)r * Npad + i];
        long long* fenv = d.f_env + (size_t)r * 3 * N;
        fx_addf(&fenv[oi], fx, (float)FORCE_SCALE);
        fx_addf(&fenv[N + oi], fy, (float)FORCE_SCALE);
        fx_addf(&fenv[2 * N + oi], fz, (float)FORCE_SCALE);
    }
    if (ENERGY) {
        // every pair appears in both atoms' lists
        const float e = warp_sum(etot);
        if (lane == 0 && e != 0.f) fx_add(&d.eacc[r * N_ETERMS + E_PAIR], 0.5 * (double)e, ENERGY_SCALE);
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_neighbor_pairs: enumerate (for tests) the non-excluded pairs within the cutoff found through the Verlet list.
// ---------------------------------------------------------------------------------------------------------
template <typename IDX>
__global__ void k_neighbor_pairs(Dev d, int r, long long* codes, unsigned long long capacity, unsigned long long* n_out) {
    const int Npad = d.Npad;
    const float4* posq_s = d.posq_s + (size_t)r * Npad;
    const int* orig_s = d.orig_s + (size_t)r * Npad;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2];
    const float ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float cut2 = d.periodic ? d.cutoff2 : 3.0e38f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
        const float4 pi = posq_s[i];
        const int oi = orig_s[i];
        const IDX* list = reinterpret_cast<const IDX*>(d.nl_list) + ((size_t)r * Npad + i) * d.nl_M;
        const int cnt = d.nl_count[(size_t)r * Npad + i];
        for (int k = 0; k < cnt; ++k) {
            const int s = (int)list[k];
            const int oj = orig_s[s];
            if (oj < oi) continue;                       // full list: report each pair once
            const float4 pj = posq_s[s];
            float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            if (d.periodic) {
                dx -= bx * rintf(dx * ibx);
                dy -= by * rintf(dy * iby);
                dz -= bz * rintf(dz * ibz);
            }
            if (dx * dx + dy * dy + dz * dz < cut2) {
                unsigned long long slot = atomicAdd(n_out, 1ull);
                if (slot < capacity) codes[slot] = (long long)oi * d.N + oj;
            }
        }
    }
}
