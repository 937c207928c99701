// Rule-independent device helpers compiled ahead of time by nvcc for sm_100a (static_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace se_static {

// 256-bin population count of material ids over `n` packed-u32 cells (ids > 255 count in bin 255).
void launch_census(const unsigned* cells, size_t n, unsigned long long* counts256, cudaStream_t stream);


// *out += sum of mix(index0 + i, cells[i]) over i < n (64-bit wrap-around): sharding-independent grid checksum.
void launch_checksum(const unsigned* cells, size_t n, unsigned long long index0, unsigned long long* out, cudaStream_t stream);

}  // namespace se_static
