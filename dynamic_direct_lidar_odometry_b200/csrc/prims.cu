// Device-wide prefix sum and stable radix sort, hand-written (no library kernels on any path of this library).
//
//   scan_int          inclusive / exclusive prefix sum of 32-bit ints: tile sums -> one-block scan of the sums -> apply
//   radix_sort_pairs  stable LSD radix sort of (32-bit key, 32-bit value) pairs, 8 bits per pass: per-tile digit
//                     histograms -> scan of the digit-major histogram table -> scatter with stable in-tile ranks
//
// Used where a cloud is too large for the 16-CTA cluster sort of cluster_sort.cu (index build of submaps and other
// clouds above 131 072 points: replaces the serial kd-tree build of nanoflann, nanoflann_impl.hpp:987-1143, together
// with index.cu), by the voxel filter (pcl::VoxelGrid's sort of the voxel indices) and by the segmentation stage.
#include <algorithm>

#include "common.cuh"
#include "prims.cuh"

namespace ddlo {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

template <class T>
__device__ __forceinline__ T warp_inclusive(T v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// block-wide exclusive prefix of one value per thread (256 threads); *total receives the block sum
template <class T>
__device__ __forceinline__ T block_exclusive(T v, T* s_warp /*[8]*/, T* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T inc = warp_inclusive(v, lane);
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  T base = 0, sum = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const T t = s_warp[w];
    if (w < warp) base += t;
    sum += t;
  }
  __syncthreads();
  if (total) *total = sum;
  return base + inc - v;
}

// Prefix sum of one tile of kScanTile elements starting at `base`, continued from `offset`; returns the tile's sum.
// Global memory is read and written coalesced through s_tile; a thread scans kScanItems CONSECUTIVE elements, so that
// its running sum continues the block prefix.
template <class T, bool kInclusive>
__device__ __forceinline__ T tile_scan(const T* __restrict__ in, T* __restrict__ out, size_t n, size_t base, T offset, T* s_tile, T* s_warp) {
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const size_t i = base + (size_t)k * kScanThreads + threadIdx.x;
    s_tile[k * kScanThreads + threadIdx.x] = i < n ? in[i] : T(0);
  }
  __syncthreads();
  T v[kScanItems];
  T sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = s_tile[threadIdx.x * kScanItems + k];
    sum += v[k];
  }
  T total = 0;
  T run = offset + block_exclusive(sum, s_warp, &total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (kInclusive) run += v[k];
    s_tile[threadIdx.x * kScanItems + k] = run;
    if (!kInclusive) run += v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const size_t i = base + (size_t)k * kScanThreads + threadIdx.x;
    if (i < n) out[i] = s_tile[k * kScanThreads + threadIdx.x];
  }
  __syncthreads();
  return total;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const T* __restrict__ in, size_t n, T* __restrict__ tile_sums) {
  __shared__ T s_warp[kScanThreads / 32];
  const size_t base = (size_t)blockIdx.x * kScanTile;
  T v = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const size_t i = base + (size_t)k * kScanThreads + threadIdx.x;
    if (i < n) v += in[i];
  }
  T total = 0;
  block_exclusive(v, s_warp, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place
template <class T>
__global__ void __launch_bounds__(kScanThreads) k_scan_spine(T* __restrict__ tile_sums, int n_tiles) {
  __shared__ T s_warp[kScanThreads / 32];
  T carry = 0;
  for (int b0 = 0; b0 < n_tiles; b0 += kScanThreads) {
    const int i = b0 + threadIdx.x;
    const T v = i < n_tiles ? tile_sums[i] : T(0);
    T total = 0;
    const T ex = block_exclusive(v, s_warp, &total);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    carry += total;
  }
}

template <class T, bool kInclusive>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const T* __restrict__ in, T* __restrict__ out, size_t n, const T* __restrict__ tile_offsets) {
  __shared__ T s_warp[kScanThreads / 32];
  __shared__ T s_tile[kScanTile];
  tile_scan<T, kInclusive>(in, out, n, (size_t)blockIdx.x * kScanTile, tile_offsets[blockIdx.x], s_tile, s_warp);
}

size_t scan_temp_bytes(size_t n) { return (((n + kScanTile - 1) / kScanTile) * sizeof(unsigned long long) + 255) & ~size_t(255); }

template <class T>
static int scan_any(cudaStream_t st, const T* in, T* out, size_t n, bool inclusive, void* temp, long long* launches) {
  if (n == 0) return DDLO_OK;
  const size_t n_tiles = (n + kScanTile - 1) / kScanTile;
  if (n_tiles > (size_t)1 << 30) return fail(DDLO_E_UNSUPPORTED, "scan: too many elements");
  T* tile_sums = static_cast<T*>(temp);
  k_scan_tiles<T><<<(unsigned)n_tiles, kScanThreads, 0, st>>>(in, n, tile_sums);
  k_scan_spine<T><<<1, kScanThreads, 0, st>>>(tile_sums, (int)n_tiles);
  if (inclusive)
    k_scan_apply<T, true><<<(unsigned)n_tiles, kScanThreads, 0, st>>>(in, out, n, tile_sums);
  else
    k_scan_apply<T, false><<<(unsigned)n_tiles, kScanThreads, 0, st>>>(in, out, n, tile_sums);
  if (launches) *launches += 3;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}
int scan_int(cudaStream_t st, const int* in, int* out, size_t n, bool inclusive, void* temp, long long* launches) {
  return scan_any<int>(st, in, out, n, inclusive, temp, launches);
}
int scan_u64(cudaStream_t st, const unsigned long long* in, unsigned long long* out, size_t n, bool inclusive, void* temp, long long* launches) {
  return scan_any<unsigned long long>(st, in, out, n, inclusive, temp, launches);
}

// ---- radix sort ---------------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsMaxRounds = 16;  // a tile is `rounds` x 256 consecutive elements
constexpr int kRsBins = 256;

// tile size for n elements: about four blocks per SM of a B200 in flight, 2..16 rounds of 256 elements
static inline int rs_rounds(int n) { return std::max(2, std::min(kRsMaxRounds, (n + 256 * 592 - 1) / (256 * 592))); }

// histogram table, digit-major: hist[digit * n_tiles + tile]
__global__ void __launch_bounds__(kRsThreads) k_rs_hist(const unsigned* __restrict__ keys, int n, int shift, int n_tiles, int rounds, int* __restrict__ hist) {
  __shared__ int s_hist[kRsBins];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * rounds * kRsThreads;
  for (int k = 0; k < rounds; ++k) {
    const int i = base + k * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&s_hist[(keys[i] >> shift) & (kRsBins - 1)], 1);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_hist[threadIdx.x];
}

// Exclusive scan of the histogram table in ONE launch: block b scans slice b (kScanTile entries) in place; the last
// block to finish scans the slice totals into slice_off, which the consumer adds (bases[j] + slice_off[j / kScanTile]).
// `zero` (the next pass's table, filled by this pass's scatter) is cleared on the way.  ctl[0]: ticket, left at zero.
__global__ void __launch_bounds__(kScanThreads) k_rs_scan_table(int* __restrict__ table, int n_entries, int* __restrict__ slice_off, int* __restrict__ ctl,
                                                                int* __restrict__ zero) {
  __shared__ int s_warp[kScanThreads / 32];
  __shared__ int s_tile[kScanTile];
  __shared__ int s_last;
  const size_t base = (size_t)blockIdx.x * kScanTile;
  const int total = tile_scan<int, false>(table, table, (size_t)n_entries, base, 0, s_tile, s_warp);
  if (zero)
    for (int k = threadIdx.x; k < kScanTile; k += kScanThreads)
      if (base + k < (size_t)n_entries) zero[base + k] = 0;
  if (threadIdx.x == 0) slice_off[blockIdx.x] = total;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ctl, 1) == (int)gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  int carry = 0;
  for (int b0 = 0; b0 < (int)gridDim.x; b0 += kScanThreads) {
    const int i = b0 + threadIdx.x;
    const int v = i < (int)gridDim.x ? __ldcg(slice_off + i) : 0;
    int tot = 0;
    const int ex = block_exclusive(v, s_warp, &tot);
    if (i < (int)gridDim.x) slice_off[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *ctl = 0;
}

// The tile's elements are taken in rounds of 256 consecutive elements (round r, thread t: element r * 256 + t of the
// tile), which is their order in the input.  The rank of an element among the tile's elements with the same digit is
//   (elements of that digit in earlier rounds) + (in lower warps of this round) + (in lower lanes of this warp),
// so equal digits keep their input order: the pass is stable.  While it places an element the kernel also counts it
// into the NEXT pass's histogram table (digit of the next pass, tile of its new position), so only the first pass
// needs a histogram kernel.
__global__ void __launch_bounds__(kRsThreads) k_rs_scatter(const unsigned* __restrict__ keys, const int* __restrict__ vals, int n, int shift,
                                                           int n_tiles, int rounds, const int* __restrict__ bases /* slice-wise exclusive scan of hist */,
                                                           const int* __restrict__ slice_off, unsigned* __restrict__ keys_out, int* __restrict__ vals_out, int* __restrict__ hist_next,
                                                           int shift_next) {
  __shared__ int s_run[kRsBins];                    // global position of the next element of each digit
  __shared__ int s_wcount[kRsThreads / 32][kRsBins];  // elements of each digit per warp in the current round
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const size_t j = (size_t)threadIdx.x * n_tiles + blockIdx.x;
    s_run[threadIdx.x] = bases[j] + slice_off[j / kScanTile];
  }
#pragma unroll
  for (int w = 0; w < kRsThreads / 32; ++w) s_wcount[w][threadIdx.x] = 0;
  __syncthreads();
  const int tile = rounds * kRsThreads;
  const int base = blockIdx.x * tile;
  for (int r = 0; r < rounds; ++r) {
    if (base + r * kRsThreads >= n) break;  // uniform
    const int i = base + r * kRsThreads + threadIdx.x;
    const bool live = i < n;
    unsigned key = 0;
    int val = 0;
    if (live) {
      key = keys[i];
      val = vals[i];
    }
    const int d = live ? (int)((key >> shift) & (kRsBins - 1)) : kRsBins + lane;  // dead lanes: a digit of their own
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int lane_rank = __popc(peers & ((1u << lane) - 1u));
    if (live && lane_rank == 0) s_wcount[warp][d] = __popc(peers);
    __syncthreads();
    if (live) {
      int pos = s_run[d] + lane_rank;
      for (int w = 0; w < warp; ++w) pos += s_wcount[w][d];
      DDLO_CHECK_INDEX(pos, n, "k_rs_scatter: output position");
      keys_out[pos] = key;
      vals_out[pos] = val;
      if (hist_next) atomicAdd(hist_next + (size_t)((key >> shift_next) & (kRsBins - 1)) * n_tiles + pos / tile, 1);
    }
    __syncthreads();
    {
      int add = 0;
#pragma unroll
      for (int w = 0; w < kRsThreads / 32; ++w) {
        add += s_wcount[w][threadIdx.x];
        s_wcount[w][threadIdx.x] = 0;
      }
      s_run[threadIdx.x] += add;
    }
    __syncthreads();
  }
}

static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

size_t radix_sort_temp_bytes(int n) {
  const int nn = std::max(n, 1);
  const size_t n_tiles = ((size_t)nn + (size_t)rs_rounds(nn) * kRsThreads - 1) / ((size_t)rs_rounds(nn) * kRsThreads);
  const size_t table = n_tiles * kRsBins;
  return 2 * al256(table * sizeof(int)) + al256(((table + kScanTile - 1) / kScanTile) * sizeof(int)) + 256;
}

int radix_sort_pairs(cudaStream_t st, unsigned* keys, unsigned* keys_alt, int* vals, int* vals_alt, int n, int bits, void* temp,
                     unsigned** keys_sorted, int** vals_sorted, long long* launches) {
  *keys_sorted = keys;
  *vals_sorted = vals;
  if (n <= 1 || bits <= 0) return DDLO_OK;
  const int rounds = rs_rounds(n);
  const int n_tiles = (n + rounds * kRsThreads - 1) / (rounds * kRsThreads);
  const size_t table = (size_t)n_tiles * kRsBins;
  int* hist_a = static_cast<int*>(temp);
  int* hist_b = reinterpret_cast<int*>(static_cast<char*>(temp) + al256(table * sizeof(int)));
  const int n_slices = (int)((table + kScanTile - 1) / kScanTile);
  int* slice_off = reinterpret_cast<int*>(static_cast<char*>(temp) + 2 * al256(table * sizeof(int)));
  int* ctl = reinterpret_cast<int*>(static_cast<char*>(temp) + 2 * al256(table * sizeof(int)) + al256((size_t)n_slices * sizeof(int)));
  DDLO_CUDA(cudaMemsetAsync(ctl, 0, sizeof(int), st));
  unsigned *kin = keys, *kout = keys_alt;
  int *vin = vals, *vout = vals_alt;
  k_rs_hist<<<n_tiles, kRsThreads, 0, st>>>(kin, n, 0, n_tiles, rounds, hist_a);
  if (launches) *launches += 1;
  for (int shift = 0; shift < bits; shift += 8) {
    const bool more = shift + 8 < bits;
    k_rs_scan_table<<<n_slices, kScanThreads, 0, st>>>(hist_a, (int)table, slice_off, ctl, more ? hist_b : nullptr);
    k_rs_scatter<<<n_tiles, kRsThreads, 0, st>>>(kin, vin, n, shift, n_tiles, rounds, hist_a, slice_off, kout, vout, more ? hist_b : nullptr, shift + 8);
    if (launches) *launches += 2;
    std::swap(kin, kout);
    std::swap(vin, vout);
    std::swap(hist_a, hist_b);
  }
  DDLO_CUDA(cudaGetLastError());
  *keys_sorted = kin;
  *vals_sorted = vin;
  return DDLO_OK;
}

}  // namespace ddlo
