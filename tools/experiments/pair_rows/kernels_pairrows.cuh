// Pair rows: one Verlet row per PAIR of adjacent cell-sorted atoms, and the packed-FP32 pair kernel that consumes it.
//
// ncu on k_pair4 at eight walkers (profiles/r02_k_pair4_8walkers_ncu_keys.txt): L1 data pipe 71 %, heavy FMA pipe 57 %,
// issue 54 % — the kernel waits for the 16-byte and 8-byte gathers of 32 different neighbours per warp instruction, not
// for arithmetic.  Two adjacent sorted atoms share most of their neighbours (union = 1.13 x one list at the T4L density),
// and Blackwell's packed FP32 operations take a broadcast operand for free: with the pair (i0, i1) in the two halves of
// every operand, one gather of j serves two interactions and no register ever has to be packed or duplicated.
//
//   row of the pair starting at sorted index s (the first, third, ... atom of a builder group): sorted indices of the atoms
//   within cutoff + skin of EITHER atom and excluded from NEITHER (stored where atom s's row used to be, same capacity);
//   the interactions that exist for only one of the two atoms (the partner is bonded to the other one, or IS the other one)
//   go to a per-walker list of singles (i, j), appended through an atomic cursor — each single is converted to fixed point
//   and added on its own, so the order of the list does not reach any sum (results stay bitwise reproducible).
//
// Full-list semantics as before: the pair (i, j) is evaluated from i's side and from j's side, forces only on the row's own
// atoms, energies halved.  PME systems with the polynomial Ewald kernel; everything else keeps per-atom rows.
#pragma once
#include "kernels_nb.cuh"

#define PR_SLACK 4           /* a chunk appends at most 4 entries per lane between two capacity checks */
#define PR_CAND_F4 40        /* 32 staged candidates in 8 pieces of 4, one float4 of padding per piece */
__host__ __device__ inline int pr_sub_stride(int cq, int idx_bytes) {
    const int words = ((cq + PR_SLACK) * idx_bytes + 3) / 4 | 1;
    return words * 4 / idx_bytes;
}
#define PR_HEAD_F4 (PR_CAND_F4 + BUILD_MAX_RUNS + (BUILD_MAX_RUNS + 4) / 4)
__host__ __device__ inline size_t pr_smem_bytes(int cq, int idx_bytes) {
    return PR_HEAD_F4 * sizeof(float4) + (size_t)32 * pr_sub_stride(cq, idx_bytes) * idx_bytes;
}

__device__ __forceinline__ void pr_append_single(const Dev& d, int r, int si, int sj) {
    Globals& g = d.g[r];
    const int slot = atomicAdd(&g.n_singles, 1);
    if (slot < d.single_cap) d.singles[(size_t)r * d.single_cap + slot] = make_int2(si, sj);
    else g.item_overflow = 1;
}

// lane = (pair p = lane & 3, candidate subset q = lane >> 2); the row of pair p is the concatenation, flush after flush, of
// the sub-lists of its eight lanes in subset order.  `done` = entries already in the row of this lane's pair.
template <typename IDX>
__device__ __noinline__ int pr_flush(const IDX* subs, int stride, IDX* rows, int nl_M, int lane, int cnt, int done) {
    __syncwarp();
    int off = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const int c = __shfl_up_sync(0xffffffffu, cnt, 4 * k);
        if (lane >= 4 * k) off += c;
    }
    const int total = __shfl_sync(0xffffffffu, off + cnt, 28 + (lane & 3));      // lane (p, 7) knows it
    const int base = done + off;
    for (int L = 0; L < 32; ++L) {
        const int nL = __shfl_sync(0xffffffffu, cnt, L), oL = __shfl_sync(0xffffffffu, base, L);
        const IDX* src = subs + (size_t)L * stride;
        IDX* dst = rows + (size_t)(2 * (L & 3)) * nl_M + oL;
        for (int k = lane; k < nL; k += 32)
            if (oL + k < nl_M) dst[k] = src[k];
    }
    __syncwarp();
    return done + total;
}

template <bool RINT, typename IDX>
__device__ __forceinline__ void pr_stream(const Dev& d, int r, const float4* __restrict__ posq_s, const int* __restrict__ orig_s,
                                          const float4* runs, const int* off, int nruns, bool rx, bool ry, bool rz, float4* cand,
                                          IDX* mysub, int cq, int lane, float4 pa0, float4 pa1, int oi0, int oi1, ull wi0, ull wi1,
                                          bool far0, bool far1, bool valid0, bool valid1, int si0, int si1, bool anyfar,
                                          const int (&og)[BUILD_GROUP], const unsigned int (&osp)[BUILD_GROUP], int& cnt, int& done,
                                          const IDX* subs, int stride, IDX* rows) {
    const float cut2 = d.list_cutoff2;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2], ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float qnan = __int_as_float(0x7fc00000);
    const int q = lane >> 2;
    const unsigned int sub_addr = (unsigned int)__cvta_generic_to_shared(mysub);
    const int total = off[nruns];
    int kp = 0;
    float4 cnext = make_float4(qnan, qnan, qnan, 0.f);
    int ojnext = 0;
    auto fetch = [&](int c) {
        cnext = make_float4(qnan, qnan, qnan, 0.f);
        if (c < total) {
            while (c >= off[kp + 1]) ++kp;
            const float4 rn = runs[kp];
            const int s = __float_as_int(rn.w) + (c - off[kp]);
            const float4 p = posq_s[s];
            ojnext = orig_s[s];
            cnext = make_float4(p.x + rn.x, p.y + rn.y, p.z + rn.z, __int_as_float(s));   // shifted image; w = sorted index
        }
    };
    auto dist2 = [&](const float4& c, const float4& p) {
        float dx = c.x - p.x, dy = c.y - p.y, dz = c.z - p.z;
        if (RINT) {
            if (rx) dx -= bx * rintf(dx * ibx);
            if (ry) dy -= by * rintf(dy * iby);
            if (rz) dz -= bz * rintf(dz * ibz);
        }
        return dx * dx + dy * dy + dz * dz;
    };
    fetch(lane);
    for (int c0 = 0; c0 < total; c0 += 32) {
        bool near = false;
        {
            const float4 c = cnext;
            const int oj = ojnext;
            const bool have = c0 + lane < total;
            fetch(c0 + 32 + lane);
            if (have) {
#pragma unroll
                for (int k = 0; k < BUILD_GROUP; ++k) near = near || (unsigned int)(oj - og[k]) <= osp[k];
            }
            __syncwarp();
            cand[lane + (lane >> 2)] = c;                           // pieces of 4, padded: conflict-free LDS.128
            __syncwarp();
        }
        const bool check = anyfar || __any_sync(0xffffffffu, near);   // warp-uniform
        unsigned int wp = sub_addr + (unsigned int)cnt * (unsigned int)sizeof(IDX);
        if (!check) {
            // no candidate of this chunk is bonded to (or is) an atom of the group: in range of either atom = entry of the row
            float r2v[4];
            int sv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float4 c = cand[q * 5 + t];
                r2v[t] = fminf(dist2(c, pa0), dist2(c, pa1));       // fminf drops the NaN of an empty second slot
                sv[t] = __float_as_int(c.w);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                sts_idx(wp, (IDX)sv[t]);
                wp += r2v[t] < cut2 ? (unsigned int)sizeof(IDX) : 0u;   // NaN (padding) compares false
            }
        } else {
#pragma unroll 1
            for (int t = 0; t < 4; ++t) {
                const float4 c = cand[q * 5 + t];
                const bool in0 = dist2(c, pa0) < cut2, in1 = dist2(c, pa1) < cut2;
                if (!(in0 || in1)) continue;
                const int sj = __float_as_int(c.w);
                const int oj = orig_s[sj];
                bool ok0 = true, ok1 = true;                           // an empty slot excludes nothing (and is never in range)
                if (valid0) {
                    const unsigned int dd = (unsigned int)(oj - oi0 + 32);
                    if (dd < 64u) ok0 = !((wi0 >> dd) & 1ull);         // includes the atom itself (bit 32)
                    else if (far0) ok0 = !pair_excluded(d, oi0, wi0, true, oj, d.has_far[oj]);
                }
                if (valid1) {
                    const unsigned int dd = (unsigned int)(oj - oi1 + 32);
                    if (dd < 64u) ok1 = !((wi1 >> dd) & 1ull);
                    else if (far1) ok1 = !pair_excluded(d, oi1, wi1, true, oj, d.has_far[oj]);
                }
                if (ok0 && ok1) { sts_idx(wp, (IDX)sj); wp += (unsigned int)sizeof(IDX); }
                else {
                    if (ok0 && in0) pr_append_single(d, r, si0, sj);
                    if (ok1 && in1) pr_append_single(d, r, si1, sj);
                }
            }
        }
        cnt = (int)((wp - sub_addr) / (unsigned int)sizeof(IDX));
        if (__any_sync(0xffffffffu, cnt > cq)) { done = pr_flush<IDX>(subs, stride, rows, d.nl_M, lane, cnt, done); cnt = 0; }
    }
}

template <typename IDX>
__global__ void __launch_bounds__(32) k_build_pairs(Dev d, int cq) {
    extern __shared__ float4 s_build[];
    float4* cand = s_build;
    const int stride = pr_sub_stride(cq, (int)sizeof(IDX));
    float4* runs = s_build + PR_CAND_F4;
    int* off = reinterpret_cast<int*>(s_build + PR_CAND_F4 + BUILD_MAX_RUNS);
    IDX* subs = reinterpret_cast<IDX*>(s_build + PR_HEAD_F4);
    const int lane = threadIdx.x;
    IDX* mysub = subs + (size_t)lane * stride;
    const int N = d.N, Npad = d.Npad;
    const float qnan = __int_as_float(0x7fc00000);
    for (int wk = 0; wk < d.R; ++wk) {
    const int r = (blockIdx.x + wk) % d.R;
    Globals& g = d.g[r];
    if (!g.do_rebuild) continue;
    const float4* __restrict__ posq_s = d.posq_s + (size_t)r * Npad;
    const int* __restrict__ orig_s = d.orig_s + (size_t)r * Npad;
    const int* __restrict__ start = d.cell_start + (size_t)r * (d.ncells + 1);
    const int* __restrict__ groups = d.group_first + (size_t)r * d.group_capacity;
    const int n_groups = g.n_groups;
    const int p = lane & 3;
    for (;;) {
        int gi = 0;
        if (lane == 0) gi = atomicAdd(&g.build_cursor, 1);
        gi = __shfl_sync(0xffffffffu, gi, 0);
        if (gi >= n_groups) break;
        const int packed = groups[gi];
        const int i0 = packed >> 4, na = packed & 15;
        const bool valid0 = 2 * p < na, valid1 = 2 * p + 1 < na;
        const int si0 = valid0 ? i0 + 2 * p : i0, si1 = valid1 ? i0 + 2 * p + 1 : i0;
        const float4 pr0 = posq_s[si0], pr1 = posq_s[si1];            // empty slots hold the first atom (bounding box only)
        const float4 pa0 = valid0 ? pr0 : make_float4(qnan, qnan, qnan, 0.f);
        const float4 pa1 = valid1 ? pr1 : make_float4(qnan, qnan, qnan, 0.f);
        const int oi0 = orig_s[si0], oi1 = orig_s[si1];
        const ull wi0 = valid0 ? (d.excl_win[oi0] | (1ull << 32)) : 0ull, wi1 = valid1 ? (d.excl_win[oi1] | (1ull << 32)) : 0ull;
        const bool far0 = valid0 ? d.has_far[oi0] : false, far1 = valid1 ? d.has_far[oi1] : false;
        const bool anyfar = __any_sync(0xffffffffu, far0 || far1);
        // exclusion window of every atom of the group as [og, og + osp] in topology indices; atom k sits in lane k >> 1
        int og[BUILD_GROUP];
        unsigned int osp[BUILD_GROUP];
        {
            const int below0 = valid0 ? 32 - (__ffsll((long long)wi0) - 1) : 0, above0 = valid0 ? 31 - __clzll((long long)wi0) : 0;
            const int below1 = valid1 ? 32 - (__ffsll((long long)wi1) - 1) : 0, above1 = valid1 ? 31 - __clzll((long long)wi1) : 0;
            const int g0 = valid0 ? oi0 - below0 : 0x3fffffff, g1 = valid1 ? oi1 - below1 : 0x3fffffff;
            const int s0 = valid0 ? below0 + above0 : 0, s1 = valid1 ? below1 + above1 : 0;
#pragma unroll
            for (int k = 0; k < BUILD_GROUP; ++k) {
                const int a = __shfl_sync(0xffffffffu, (k & 1) ? g1 : g0, k >> 1);
                const int b = __shfl_sync(0xffffffffu, (k & 1) ? s1 : s0, k >> 1);
                og[k] = a; osp[k] = (unsigned int)b;
            }
        }
        int cnt = 0, done = 0;
        IDX* rows = reinterpret_cast<IDX*>(d.nl_list) + ((size_t)r * Npad + i0) * d.nl_M;
        if (!d.periodic) {
            if (lane == 0) { runs[0] = make_float4(0.f, 0.f, 0.f, __int_as_float(0)); off[0] = 0; off[1] = N; }
            __syncwarp();
            pr_stream<false, IDX>(d, r, posq_s, orig_s, runs, off, 1, false, false, false, cand, mysub, cq, lane, pa0, pa1, oi0,
                                  oi1, wi0, wi1, far0, far1, valid0, valid1, si0, si1, anyfar, og, osp, cnt, done, subs, stride, rows);
        } else {
            const int ncx = d.ncell[0], ncy = d.ncell[1], ncz = d.ncell[2];
            int cx0, cy0, cz0, cx1, cy1, cz1;
            atom_cell_coords(d, pr0, cx0, cy0, cz0);
            atom_cell_coords(d, pr1, cx1, cy1, cz1);
            const int xa = __reduce_min_sync(0xffffffffu, min(cx0, cx1)), xb = __reduce_max_sync(0xffffffffu, max(cx0, cx1));
            const int ya = __reduce_min_sync(0xffffffffu, min(cy0, cy1)), yb = __reduce_max_sync(0xffffffffu, max(cy0, cy1));
            const int za = __reduce_min_sync(0xffffffffu, min(cz0, cz1)), zb = __reduce_max_sync(0xffffffffu, max(cz0, cz1));
            const float lox = warp_min(fminf(pr0.x, pr1.x)), hix = warp_max(fmaxf(pr0.x, pr1.x));
            const float loy = warp_min(fminf(pr0.y, pr1.y)), hiy = warp_max(fmaxf(pr0.y, pr1.y));
            const float loz = warp_min(fminf(pr0.z, pr1.z)), hiz = warp_max(fmaxf(pr0.z, pr1.z));
            const bool rx = xb - xa + 5 > ncx, ry = yb - ya + 5 > ncy, rz = zb - za + 5 > ncz;
            const int x0 = rx ? 0 : xa - 2, x1 = rx ? ncx - 1 : xb + 2;
            const int y0 = ry ? 0 : ya - 2, y1 = ry ? ncy - 1 : yb + 2;
            const int nruns = build_group_runs(d, start, lane, rx, ry, rz, x0, x1, y0, y1, za, zb, lox, hix, loy, hiy, loz,
                                               hiz, runs, off);
            if (rx || ry || rz)
                pr_stream<true, IDX>(d, r, posq_s, orig_s, runs, off, nruns, rx, ry, rz, cand, mysub, cq, lane, pa0, pa1, oi0, oi1,
                                     wi0, wi1, far0, far1, valid0, valid1, si0, si1, anyfar, og, osp, cnt, done, subs, stride, rows);
            else
                pr_stream<false, IDX>(d, r, posq_s, orig_s, runs, off, nruns, false, false, false, cand, mysub, cq, lane, pa0, pa1,
                                      oi0, oi1, wi0, wi1, far0, far1, valid0, valid1, si0, si1, anyfar, og, osp, cnt, done, subs,
                                      stride, rows);
        }
        const int total = pr_flush<IDX>(subs, stride, rows, d.nl_M, lane, cnt, done);
        if (__any_sync(0xffffffffu, total > d.nl_M)) { if (lane == 0) g.item_overflow = 1; }
        if (lane < 4) {                                       // lanes (p, 0): the counts of pair p; the odd rows stay empty
            if (valid0) d.nl_count[(size_t)r * Npad + i0 + 2 * p] = min(total, d.nl_M);
            if (valid1) d.nl_count[(size_t)r * Npad + i0 + 2 * p + 1] = 0;
        }
        __syncwarp();
    }
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_pair5: LJ + polynomial-Ewald direct-space forces (and, on request, energies) over the pair rows.  One warp per builder
// group: lane = (pair p = lane >> 3, part = lane & 7); the eight lanes of a pair stride over its row like k_pair4's lanes
// over an atom's row, with the same two-deep software pipeline; every packed operand holds (atom 0, atom 1) of the pair
// and takes the neighbour as its broadcast operand.  The CTAs behind the rows evaluate the singles.
// ---------------------------------------------------------------------------------------------------------
struct PairAcc6 { f32x2 fx, fy, fz; };

template <int DEG, bool ENERGY>
__device__ __forceinline__ void pair_slot2(const Dev& d, const float4 pj, const float2 ej, bool inrow, f32x2 pix, f32x2 piy,
                                           f32x2 piz, f32x2 sx, f32x2 ey, f32x2 qi2, float cut2, PairAcc6& acc, float& etot) {
    const f32x2 magic = f2_dup(12582912.0f), nmagic = f2_dup(-12582912.0f);
    // d = r_j - r_i (pix.. hold the NEGATED coordinates of the pair): the broadcast operand needs no negation, and the
    // accumulators collect -F, negated once at the end
    f32x2 dx = f2_add(pix, f2_dup(pj.x));
    f32x2 dy = f2_add(piy, f2_dup(pj.y));
    f32x2 dz = f2_add(piz, f2_dup(pj.z));
    dx = f2_fma(f2_dup(-d.boxf[0]), f2_add(f2_fma(dx, f2_dup(d.boxf[3]), magic), nmagic), dx);
    dy = f2_fma(f2_dup(-d.boxf[1]), f2_add(f2_fma(dy, f2_dup(d.boxf[4]), magic), nmagic), dy);
    dz = f2_fma(f2_dup(-d.boxf[2]), f2_add(f2_fma(dz, f2_dup(d.boxf[5]), magic), nmagic), dz);
    const f32x2 r2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
    float r2a, r2b;
    f2_unpack(r2, r2a, r2b);
    const bool ina = inrow && (r2a < cut2), inb = inrow && (r2b < cut2);
    const float ira = rsqrtf(r2a), irb = rsqrtf(r2b);
    const f32x2 invr = f2_pack(ira, irb);
    const f32x2 invr2 = f2_mul(invr, invr);
    const f32x2 sig = f2_add(sx, f2_dup(ej.x));
    const f32x2 s2 = f2_mul(f2_mul(sig, sig), invr2);
    const f32x2 s6 = f2_mul(f2_mul(s2, s2), s2);
    const f32x2 eps4 = f2_mul(ey, f2_dup(ej.y));
    f32x2 de = f2_mul(f2_mul(eps4, f2_mul(s6, f2_fma(s6, f2_dup(12.0f), f2_dup(-6.0f)))), invr2);
    const f32x2 qq = f2_mul(qi2, f2_dup(pj.w));
    const f32x2 tt = f2_fma(r2, f2_dup(d.ewk_scale), f2_dup(-1.0f));
    f32x2 k = f2_dup(d.ewk2[DEG]);
#pragma unroll
    for (int c = DEG - 1; c >= 0; --c) k = f2_fma(k, tt, f2_dup(d.ewk2[c]));
    de = f2_fma(qq, f2_fma(f2_dup(-d.alpha3), k, f2_mul(invr, invr2)), de);
    float dea, deb;
    f2_unpack(de, dea, deb);
    de = f2_pack(ina ? dea : 0.f, inb ? deb : 0.f);
    acc.fx = f2_fma(dx, de, acc.fx);
    acc.fy = f2_fma(dy, de, acc.fy);
    acc.fz = f2_fma(dz, de, acc.fz);
    if (ENERGY) {
        float e4a, e4b, s6a, s6b, qqa, qqb;
        f2_unpack(eps4, e4a, e4b); f2_unpack(s6, s6a, s6b); f2_unpack(qq, qqa, qqb);
        if (ina) { const float ar = d.alpha * r2a * ira; etot += e4a * (s6a * s6a - s6a) + qqa * ira * erfc_times(ar, __expf(-ar * ar)); }
        if (inb) { const float ar = d.alpha * r2b * irb; etot += e4b * (s6b * s6b - s6b) + qqb * irb * erfc_times(ar, __expf(-ar * ar)); }
    }
}

template <typename IDX, int U, int DEG, bool ENERGY>
__global__ void __launch_bounds__(NL_BLOCK) k_pair5(Dev d, int skip_frozen, int n_pair_blocks) {
    const int r = blockIdx.y;
    const int N = d.N, Npad = d.Npad;
    const Globals& g = d.g[r];
    const float4* __restrict__ posq_s = d.posq_s + (size_t)r * Npad;
    const float2* __restrict__ sigeps_s = d.sigeps_s + (size_t)r * Npad;
    const float cut2 = d.cutoff2;
    long long* fenv = d.f_env + (size_t)r * 3 * N;
    if ((int)blockIdx.x >= n_pair_blocks) {
        // ---- singles: interaction of atom i with atom j that the partner of i does not have (scalar arithmetic)
        const int n = min(g.n_singles, d.single_cap);
        const int2* sg = d.singles + (size_t)r * d.single_cap;
        for (int e = ((int)blockIdx.x - n_pair_blocks) * NL_BLOCK + threadIdx.x; e < n; e += ((int)gridDim.x - n_pair_blocks) * NL_BLOCK) {
            const int2 ij = sg[e];
            if (skip_frozen && !d.mobile_s[(size_t)r * Npad + ij.x]) continue;
            const float4 pi = posq_s[ij.x], pj = posq_s[ij.y];
            const float2 si = sigeps_s[ij.x], sj = sigeps_s[ij.y];
            float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            dx -= d.boxf[0] * rint_fma(dx * d.boxf[3]);
            dy -= d.boxf[1] * rint_fma(dy * d.boxf[4]);
            dz -= d.boxf[2] * rint_fma(dz * d.boxf[5]);
            const float r2 = dx * dx + dy * dy + dz * dz;
            if (!(r2 < cut2)) continue;
            const float invr = rsqrtf(r2), invr2 = invr * invr;
            const float sig = si.x + sj.x, s2 = sig * sig * invr2, s6 = s2 * s2 * s2, eps4 = si.y * sj.y;
            const float qq = pi.w * (float)ONE_4PI_EPS0 * pj.w;
            const float tt = fmaf(r2, d.ewk_scale, -1.0f);
            float k = d.ewk2[DEG];
#pragma unroll
            for (int c = DEG - 1; c >= 0; --c) k = fmaf(k, tt, d.ewk2[c]);
            const float de = eps4 * (s6 * (12.0f * s6 - 6.0f)) * invr2 + qq * fmaf(-d.alpha3, k, invr * invr2);
            const int oi = d.orig_s[(size_t)r * Npad + ij.x];
            fx_addf(&fenv[oi], dx * de, (float)FORCE_SCALE);
            fx_addf(&fenv[N + oi], dy * de, (float)FORCE_SCALE);
            fx_addf(&fenv[2 * N + oi], dz * de, (float)FORCE_SCALE);
            if (ENERGY) {
                // added entry by entry in fixed point: the order of the singles list (an atomic cursor) must not reach a sum
                const float ar = d.alpha * r2 * invr;
                const float e = eps4 * (s6 * s6 - s6) + qq * invr * erfc_times(ar, __expf(-ar * ar));
                fx_add(&d.eacc[r * N_ETERMS + E_PAIR], 0.5 * (double)e, ENERGY_SCALE);
            }
        }
        return;
    }
    const int lane = threadIdx.x & 31;
    const int gi = (int)blockIdx.x * (NL_BLOCK / 32) + (threadIdx.x >> 5);
    if (gi >= g.n_groups) return;                                   // warp-uniform
    const int packed = d.group_first[(size_t)r * d.group_capacity + gi];
    const int i0 = packed >> 4, na = packed & 15;
    const int p = lane >> 3, part = lane & 7;
    const bool has_a = 2 * p < na, has_b = 2 * p + 1 < na;
    const int ia = has_a ? i0 + 2 * p : i0, ib = has_b ? i0 + 2 * p + 1 : i0;
    int cnt = has_a ? d.nl_count[(size_t)r * Npad + ia] : 0;
    if (skip_frozen && !(d.mobile_s[(size_t)r * Npad + ia] || (has_b && d.mobile_s[(size_t)r * Npad + ib]))) cnt = 0;
    const IDX* __restrict__ lp = reinterpret_cast<const IDX*>(d.nl_list) + ((size_t)r * Npad + ia) * d.nl_M + part;
    const int nm = cnt > part ? (cnt - part + 7) / 8 : 0;           // entries owned by this lane
    const int ntrip = (nm + U - 1) / U;
    const float qnan = __int_as_float(0x7fc00000);
    const float4 pa = posq_s[ia];
    const float4 pbr = posq_s[ib];
    const float4 pb = has_b ? pbr : make_float4(qnan, qnan, qnan, 0.f);
    const float2 sa = sigeps_s[ia], sb = sigeps_s[ib];
    const f32x2 pix = f2_pack(-pa.x, -pb.x), piy = f2_pack(-pa.y, -pb.y), piz = f2_pack(-pa.z, -pb.z);
    const f32x2 sx = f2_pack(sa.x, sb.x), ey = f2_pack(sa.y, sb.y);
    const f32x2 qi2 = f2_pack(pa.w * (float)ONE_4PI_EPS0, pb.w * (float)ONE_4PI_EPS0);
    PairAcc6 acc;
    acc.fx = acc.fy = acc.fz = f2_dup(0.f);
    float etot = 0.f;
    int iA[U], iB[U];
    float4 pA[U], pB[U];
    float2 eA[U], eB[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { iA[u] = ldg_nc_idx(lp + 8 * u); iB[u] = ldg_nc_idx(lp + 8 * (U + u)); }
#pragma unroll
    for (int u = 0; u < U; ++u) { pA[u] = ldg_nc_f4(posq_s + iA[u]); eA[u] = ldg_nc_f2(sigeps_s + iA[u]); }
#pragma unroll
    for (int u = 0; u < U; ++u) { pB[u] = ldg_nc_f4(posq_s + iB[u]); eB[u] = ldg_nc_f2(sigeps_s + iB[u]); }
#pragma unroll
    for (int u = 0; u < U; ++u) { iA[u] = ldg_nc_idx(lp + 8 * (2 * U + u)); iB[u] = ldg_nc_idx(lp + 8 * (3 * U + u)); }
    for (int t = 0; t < ntrip; t += 2) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            pair_slot2<DEG, ENERGY>(d, pA[u], eA[u], t * U + u < nm, pix, piy, piz, sx, ey, qi2, cut2, acc, etot);
#pragma unroll
        for (int u = 0; u < U; ++u) { pA[u] = ldg_nc_f4(posq_s + iA[u]); eA[u] = ldg_nc_f2(sigeps_s + iA[u]); }
#pragma unroll
        for (int u = 0; u < U; ++u) iA[u] = ldg_nc_idx(lp + 8 * ((t + 4) * U + u));
#pragma unroll
        for (int u = 0; u < U; ++u)
            pair_slot2<DEG, ENERGY>(d, pB[u], eB[u], (t + 1) * U + u < nm, pix, piy, piz, sx, ey, qi2, cut2, acc, etot);
#pragma unroll
        for (int u = 0; u < U; ++u) { pB[u] = ldg_nc_f4(posq_s + iB[u]); eB[u] = ldg_nc_f2(sigeps_s + iB[u]); }
#pragma unroll
        for (int u = 0; u < U; ++u) iB[u] = ldg_nc_idx(lp + 8 * ((t + 5) * U + u));
    }
    float fxa, fxb, fya, fyb, fza, fzb;
    f2_unpack(acc.fx, fxa, fxb); f2_unpack(acc.fy, fya, fyb); f2_unpack(acc.fz, fza, fzb);
    fxa = -fxa; fya = -fya; fza = -fza; fxb = -fxb; fyb = -fyb; fzb = -fzb;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        fxa += __shfl_xor_sync(0xffffffffu, fxa, o); fya += __shfl_xor_sync(0xffffffffu, fya, o); fza += __shfl_xor_sync(0xffffffffu, fza, o);
        fxb += __shfl_xor_sync(0xffffffffu, fxb, o); fyb += __shfl_xor_sync(0xffffffffu, fyb, o); fzb += __shfl_xor_sync(0xffffffffu, fzb, o);
    }
    if (part == 0 && has_a && cnt > 0) {
        const int oi = d.orig_s[(size_t)r * Npad + ia];
        fx_addf(&fenv[oi], fxa, (float)FORCE_SCALE);
        fx_addf(&fenv[N + oi], fya, (float)FORCE_SCALE);
        fx_addf(&fenv[2 * N + oi], fza, (float)FORCE_SCALE);
    }
    if (part == 1 && has_b && cnt > 0) {
        const int oi = d.orig_s[(size_t)r * Npad + ib];
        fx_addf(&fenv[oi], fxb, (float)FORCE_SCALE);
        fx_addf(&fenv[N + oi], fyb, (float)FORCE_SCALE);
        fx_addf(&fenv[2 * N + oi], fzb, (float)FORCE_SCALE);
    }
    if (ENERGY) {
        const float e = warp_sum(etot);
        if (lane == 0 && e != 0.f) fx_add(&d.eacc[r * N_ETERMS + E_PAIR], 0.5 * (double)e, ENERGY_SCALE);
    }
}

// enumerate (for tests) the non-excluded pairs within the cutoff found through the pair rows and the singles
template <typename IDX>
__global__ void k_neighbor_pairs_pr(Dev d, int r, long long* codes, unsigned long long capacity, unsigned long long* n_out) {
    const int Npad = d.Npad;
    const float4* posq_s = d.posq_s + (size_t)r * Npad;
    const int* orig_s = d.orig_s + (size_t)r * Npad;
    const float bx = d.boxf[0], by = d.boxf[1], bz = d.boxf[2], ibx = d.boxf[3], iby = d.boxf[4], ibz = d.boxf[5];
    const float cut2 = d.cutoff2;
    auto emit = [&](int si, int sj) {
        const int oi = orig_s[si], oj = orig_s[sj];
        if (oj < oi) return;                                       // full-list semantics: report each pair once
        const float4 pi = posq_s[si], pj = posq_s[sj];
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        dx -= bx * rintf(dx * ibx); dy -= by * rintf(dy * iby); dz -= bz * rintf(dz * ibz);
        if (dx * dx + dy * dy + dz * dz < cut2) {
            unsigned long long slot = atomicAdd(n_out, 1ull);
            if (slot < capacity) codes[slot] = (long long)oi * d.N + oj;
        }
    };
    const Globals& g = d.g[r];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int w = tid; w < g.n_groups * 4; w += nt) {
        const int packed = d.group_first[(size_t)r * d.group_capacity + (w >> 2)];
        const int i0 = packed >> 4, na = packed & 15, p = w & 3;
        if (2 * p >= na) continue;
        const int ia = i0 + 2 * p, ib = 2 * p + 1 < na ? ia + 1 : -1;
        const IDX* list = reinterpret_cast<const IDX*>(d.nl_list) + ((size_t)r * Npad + ia) * d.nl_M;
        const int cnt = d.nl_count[(size_t)r * Npad + ia];
        for (int k = 0; k < cnt; ++k) {
            const int s = (int)list[k];
            emit(ia, s);
            if (ib >= 0) emit(ib, s);
        }
    }
    const int n = min(g.n_singles, d.single_cap);
    for (int e = tid; e < n; e += nt) { const int2 ij = d.singles[(size_t)r * d.single_cap + e]; emit(ij.x, ij.y); }
}
