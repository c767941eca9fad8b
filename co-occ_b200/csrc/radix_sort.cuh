// radix_sort.cuh -- segmented LSD radix sort of (uint32 key, uint32 value) pairs (defined in occ_loss.cu):
// 10 bits per pass, one warp per 4096-item chunk, __match_any_sync ranks keep every pass stable.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace coocc {
size_t radix_counters_bytes(int n, int nseg);
int radix_sort_pairs(uint32_t* kA, uint32_t* vA, uint32_t* kB, uint32_t* vB, int n, int nseg, int bits, int* counters,
                     cudaStream_t st, uint32_t** k_out, uint32_t** v_out);
}  // namespace coocc
