// Shared device helpers: fixed-point accumulation, Philox4x32-10, warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BL_WARP 32
#define FORCE_SCALE 67108864.0          /* 2^26: forces accumulate as int64 fixed point (deterministic)  */
#define ENERGY_SCALE 16777216.0         /* 2^24: energies likewise                                        */
#define GRID_SCALE 4294967296.0         /* 2^32: PME charge grid                                          */
#define ONE_4PI_EPS0 138.935456
#define TWO_OVER_SQRT_PI 1.1283791670955126
#define PME_ORDER 5

typedef unsigned long long ull;

__device__ __forceinline__ void fx_add(long long* addr, double v, double scale) {
    atomicAdd(reinterpret_cast<ull*>(addr), static_cast<ull>(static_cast<long long>(v * scale)));
}
__device__ __forceinline__ void fx_addf(long long* addr, float v, float scale) {
    atomicAdd(reinterpret_cast<ull*>(addr), static_cast<ull>(static_cast<long long>(v * scale)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- Philox4x32-10 (Salmon et al. 2011); identical to oracle/ncmc_oracle.py::philox4x32 -----------------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 r = {c0, c1, c2, c3};
    return r;
}

#define STREAM_LANGEVIN 0u
#define STREAM_VELOCITY 1u
#define STREAM_MOVE 2u
#define STREAM_ACCEPT 3u
#define STREAM_MD 4u

__host__ __device__ __forceinline__ double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

// three standard normals for (seed, stream, replica, counter, index) — Box–Muller in double.
// Deliberately not inlined: the double-precision log / sincospi expansions are ~1k instructions and the integrator
// calls this once per atom and thermostat op; inlining them everywhere overflows the instruction cache.
__device__ __noinline__ double3 philox_normal3v(uint64_t seed, uint32_t stream, uint32_t replica, uint32_t counter,
                                                uint32_t index) {
    Philox4 r = philox4x32_10(index, counter, replica, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
    double r0 = sqrt(-2.0 * log(u01(r.x)));
    double r1 = sqrt(-2.0 * log(u01(r.z)));
    double s, c;
    sincospi(2.0 * u01(r.y), &s, &c);
    return make_double3(r0 * c, r0 * s, r1 * cospi(2.0 * u01(r.w)));
}
__device__ __forceinline__ void philox_normal3(uint64_t seed, uint32_t stream, uint32_t replica, uint32_t counter,
                                               uint32_t index, double& n0, double& n1, double& n2) {
    const double3 v = philox_normal3v(seed, stream, replica, counter, index);
    n0 = v.x; n1 = v.y; n2 = v.z;
}
