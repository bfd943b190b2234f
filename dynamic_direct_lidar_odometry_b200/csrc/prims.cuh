// Device-wide prefix sum and stable radix sort (prims.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

namespace ddlo {

// bytes of scratch `scan_int` needs for n elements
size_t scan_temp_bytes(size_t n);
// prefix sum of n ints on stream st; in == out is allowed; *launches (optional) counts the kernels launched
int scan_int(cudaStream_t st, const int* in, int* out, size_t n, bool inclusive, void* temp, long long* launches);
int scan_u64(cudaStream_t st, const unsigned long long* in, unsigned long long* out, size_t n, bool inclusive, void* temp, long long* launches);

size_t radix_sort_temp_bytes(int n);
// Stable sort of (key, value) pairs by the key bits [0, bits).  keys / vals hold the input, *_alt are buffers of the
// same size; the result ends up in one of the two, reported through keys_sorted / vals_sorted.
int radix_sort_pairs(cudaStream_t st, unsigned* keys, unsigned* keys_alt, int* vals, int* vals_alt, int n, int bits, void* temp,
                     unsigned** keys_sorted, int** vals_sorted, long long* launches);

}  // namespace ddlo
