// Rule-independent kernels (sm_100a, nvcc ahead-of-time).
#include "static_kernels.h"

namespace se_static {

// Census: HBM-bound single pass (4 B/cell).  128-bit loads, per-warp privatised shared histograms
// (8 copies) to keep shared-memory atomic contention low when one material dominates, then one
// global atomicAdd per non-empty bin per CTA.  Grid = 148 SMs x 8 CTAs, grid-stride loop.
__global__ void __launch_bounds__(256) census_kernel(const unsigned* __restrict__ cells, size_t n, unsigned long long* __restrict__ counts) {
    __shared__ unsigned hist[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&hist[0][0])[i] = 0u;
    __syncthreads();
    unsigned* my = hist[threadIdx.x >> 5];
    const size_t n4 = n / 4;
    const uint4* c4 = reinterpret_cast<const uint4*>(cells);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const uint4 v = __ldg(c4 + i);
        atomicAdd(&my[min(v.x, 255u)], 1u);
        atomicAdd(&my[min(v.y, 255u)], 1u);
        atomicAdd(&my[min(v.z, 255u)], 1u);
        atomicAdd(&my[min(v.w, 255u)], 1u);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) atomicAdd(&my[min(cells[n4 * 4 + threadIdx.x], 255u)], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < 256; b += blockDim.x) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += hist[w][b];
        if (t) atomicAdd(&counts[b], t);
    }
}

// Checksum: sum over cells of mix(global cell index, id) modulo 2^64.  A sum, so it does not depend on the order the
// cells are visited in nor on how the grid is cut into strips: the per-strip values of a sharded run add up to the value
// of the unsharded grid (bench.py's parity field; the same function in numpy: sandengine_b200/grids.py grid_checksum).
__device__ __forceinline__ unsigned long long checksum_mix(unsigned long long index, unsigned id) {
    unsigned long long h = index * 0x9E3779B97F4A7C15ull + (unsigned long long)id * 0xD6E8FEB86659FD93ull;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    return h;
}

__global__ void __launch_bounds__(256) checksum_kernel(const unsigned* __restrict__ cells, size_t n, unsigned long long index0, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += checksum_mix(index0 + i, __ldg(cells + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

void launch_checksum(const unsigned* cells, size_t n, unsigned long long index0, unsigned long long* out, cudaStream_t stream) {
    checksum_kernel<<<148 * 8, 256, 0, stream>>>(cells, n, index0, out);
}

void launch_census(const unsigned* cells, size_t n, unsigned long long* counts256, cudaStream_t stream) {
    census_kernel<<<148 * 8, 256, 0, stream>>>(cells, n, counts256);
}

}  // namespace se_static
