// Micro-benchmarks behind two design decisions of round 2 (results: profiles/r2_ubench.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench scripts/ubench.cu ; gpurun -- scripts/ubench
// 1. gather: every lane of a warp reads a 4-byte (or 8-byte) entry at an independent pseudo-random index of a table of
//    `table_mb` MB -- the access pattern of a transition table kept in global memory (rule sets with up to ~100
//    materials: N^4 entries per view).  Reports lookups per SM clock for tables that are L1-, L2- and HBM-resident.
// 2. pipes: is IDP.4A (dp4a) issued on the ALU pipe or the FMA pipe?  A loop of dp4a mixed with LOP3s (ALU) against
//    the same loop mixed with IMADs (FMA): the mix that shares a pipe runs at the sum of the two, the other at the max.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static __device__ __forceinline__ unsigned hashi(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int WORDS, int ILP, int CACHE>
__global__ void __launch_bounds__(1024, 1) gather_kernel(const unsigned* __restrict__ table, unsigned mask, int iters, unsigned* out) {
    unsigned acc = 0;
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    for (int it = 0; it < iters; ++it) {
        unsigned idx[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) { x = x * 1664525u + 1013904223u; idx[j] = (hashi(x) & mask) * WORDS; }
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (WORDS == 1) {
                unsigned v;
                if (CACHE == 0) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(table + idx[j]));
                else if (CACHE == 1) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(table + idx[j]));
                else asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(table + idx[j]));
                acc += v;
            } else {
                unsigned v, w;
                asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(w) : "l"(table + idx[j]));
                acc += v ^ w;
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) pipe_kernel(int iters, unsigned* out) {
    unsigned a = threadIdx.x, b = threadIdx.x * 3u + 1u, c = threadIdx.x * 7u + 5u, d = threadIdx.x ^ 0x55u;
    unsigned e = a + 11u, f = b + 13u, g = c + 17u, h = d + 19u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0 || MODE == 1 || MODE == 2) {   // dp4a chains
                a = __dp4a(a, 0x01020304u, b); b = __dp4a(b, 0x04030201u, c); c = __dp4a(c, 0x01010101u, d); d = __dp4a(d, 0x02020202u, a);
            }
            if (MODE == 1 || MODE == 3) {                // ALU: LOP3 (xor with shift kept out: pure LOP3)
                e = (e ^ f) & (g | 0x0f0f0f0fu); f = (f ^ g) & (h | 0x33333333u); g = (g ^ h) & (e | 0x55555555u); h = (h ^ e) & (f | 0x0000ffffu);
            }
            if (MODE == 2 || MODE == 4) {                // FMA pipe: IMAD
                e = e * 0x7feb352du + f; f = f * 0x846ca68bu + g; g = g * 0x9e3779b9u + h; h = h * 0x85ebca6bu + e;
            }
        }
    }
    if ((a ^ b ^ c ^ d ^ e ^ f ^ g ^ h) == 0x12345678u) out[0] = a;
}

template <class F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();                                   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    const int n_sm = prop.multiProcessorCount;
    printf("device %s, %d SMs, %d kHz\n", prop.name, n_sm, clock_khz);
    unsigned* out;
    CK(cudaMalloc(&out, 4));
    // ---- 1. gathers ----
    const size_t max_bytes = (size_t)512 << 20;
    unsigned* table;
    CK(cudaMalloc(&table, max_bytes));
    CK(cudaMemset(table, 1, max_bytes));
    const int iters = 256;
    for (int mb : {0, 1, 16, 64, 128, 256, 512}) {          // 0 = 64 KB (L1-resident)
        const size_t bytes = mb ? (size_t)mb << 20 : (size_t)64 << 10;
        const unsigned mask4 = (unsigned)(bytes / 4 - 1), mask8 = (unsigned)(bytes / 8 - 1);
        auto report = [&](const char* name, float ms, int ilp) {
            const double lookups = (double)n_sm * 1024 * iters * ilp;
            const double per_sm_clk = lookups / n_sm / (ms * 1e-3 * clock_khz * 1e3);
            printf("gather %-28s table %4d MB: %8.3f ms  %7.1f G lookups/s  %.3f lookups/clk/SM  (x4 cells: %.0f Gcell/s)\n", name, mb, ms,
                   lookups / ms / 1e6, per_sm_clk, 4.0 * lookups / ms / 1e6);
        };
        report("nc.u32 ilp4", time_ms([&] { gather_kernel<1, 4, 0><<<n_sm, 1024>>>(table, mask4, iters, out); }), 4);
        report("nc.u32 ilp8", time_ms([&] { gather_kernel<1, 8, 0><<<n_sm, 1024>>>(table, mask4, iters, out); }), 8);
        report("cg.u32 ilp8", time_ms([&] { gather_kernel<1, 8, 1><<<n_sm, 1024>>>(table, mask4, iters, out); }), 8);
        report("nc.no_allocate.u32 ilp8", time_ms([&] { gather_kernel<1, 8, 2><<<n_sm, 1024>>>(table, mask4, iters, out); }), 8);
        report("nc.v2.u32 ilp8", time_ms([&] { gather_kernel<2, 8, 0><<<n_sm, 1024>>>(table, mask8, iters, out); }), 8);
    }
    // ---- 2. pipes ----
    const int pit = 4096;
    const char* names[5] = {"dp4a only", "dp4a + LOP3", "dp4a + IMAD", "LOP3 only", "IMAD only"};
    float ms[5];
    ms[0] = time_ms([&] { pipe_kernel<0><<<n_sm, 1024>>>(pit, out); });
    ms[1] = time_ms([&] { pipe_kernel<1><<<n_sm, 1024>>>(pit, out); });
    ms[2] = time_ms([&] { pipe_kernel<2><<<n_sm, 1024>>>(pit, out); });
    ms[3] = time_ms([&] { pipe_kernel<3><<<n_sm, 1024>>>(pit, out); });
    ms[4] = time_ms([&] { pipe_kernel<4><<<n_sm, 1024>>>(pit, out); });
    for (int k = 0; k < 5; ++k) printf("pipe %-14s %8.3f ms\n", names[k], ms[k]);
    printf("reading: if (dp4a + LOP3) ~ dp4a + LOP3 alone added up, dp4a shares the ALU pipe; if ~ max of the two, it does not (same for IMAD)\n");
    return 0;
}
