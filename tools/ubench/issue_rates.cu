// Microbenchmark (diagnostic, not product): per-SM issue rates of the instruction kinds the pair kernel is made of, on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/issue_rates tools/ubench/issue_rates.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long ull;
__device__ __forceinline__ ull pk(float a, float b) { ull r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(ull v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ ull fma2(ull a, ull b, ull c) { ull r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

#define ITERS 2048
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed) {
    float a[16];
    ull p[8];
    int n[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = pk(a[2 * i], a[2 * i + 1]); n[i] = threadIdx.x + i; }
    const float b = 1.0001f, c = 0.5f;
    const ull b2 = pk(b, b), c2 = pk(c, c);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {            // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
        } else if (MODE == 1) {     // 8 FFMA2 (= 16 lanes-FMA)
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], b2, c2);
        } else if (MODE == 2) {     // 8 FFMA2 + 8 integer ops
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], b2, c2); n[i] = (n[i] ^ it) + 3 * n[i]; }
        } else if (MODE == 3) {     // 16 FFMA + 8 integer ops
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
#pragma unroll
            for (int i = 0; i < 8; ++i) n[i] = (n[i] ^ it) + 3 * n[i];
        } else if (MODE == 4) {     // 16 MUFU.RSQ
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = rsqrtf(a[i]);
        } else if (MODE == 5) {     // 16 SHFL
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3));
        } else if (MODE == 6) {     // 8 REDUX
#pragma unroll
            for (int i = 0; i < 8; ++i) n[i] = __reduce_add_sync(0xffffffffu, n[i]) + it;
        } else if (MODE == 7) {     // 8 FFMA2 + 4 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], b2, c2);
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = rsqrtf(a[i]);
        } else if (MODE == 8) {     // 16 FFMA + 4 MUFU
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = rsqrtf(a[i]);
        } else if (MODE == 9) {     // 16 FSETP+FSEL style selects
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] < seed ? a[i] + 1.0f : 0.f;
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; upk(p[i], x, y); s += x + y + n[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int instr_per_iter, int ctas_per_sm) {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm * ctas_per_sm;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * grid * 256);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    k<MODE><<<grid, 256>>>(out, cyc, 1.5f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, cyc, 1.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)malloc(sizeof(long long) * grid);
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    const double warp_instr_per_sm = (double)ctas_per_sm * 8 * ITERS * instr_per_iter;
    printf("%-28s ctas/SM %d: %.3f warp-instr/clk/SM (in-kernel clock), %.3f ms, %.2f Gwarp-instr/s chip\n", name, ctas_per_sm,
           warp_instr_per_sm / mean, ms, warp_instr_per_sm * nsm / (ms * 1e-3) * 1e-9);
    cudaFree(out); cudaFree(cyc); free(h);
}
int main() {
    for (int c : {4, 8}) {
        run<0>("16 FFMA", 16, c);
        run<1>("8 FFMA2", 8, c);
        run<2>("8 FFMA2 + 8x(LOP,IMAD)", 24, c);
        run<3>("16 FFMA + 8x(LOP,IMAD)", 32, c);
        run<4>("16 MUFU.RSQ", 16, c);
        run<5>("16 SHFL", 16, c);
        run<6>("8 REDUX(+IADD)", 16, c);
        run<7>("8 FFMA2 + 4 MUFU", 12, c);
        run<8>("16 FFMA + 4 MUFU", 20, c);
        run<9>("16 x (FSETP,FADD,FSEL)", 48, c);
    }
    return 0;
}
