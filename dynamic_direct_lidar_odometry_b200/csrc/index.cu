// kNN index construction on the device: replaces the serial kd-tree build of
// KdTreeFLANN::setInputCloud -> KDTreeSingleIndexAdaptor::buildIndex / divideTree
// (R/nanoflann.hpp:137-143, R/impl/nanoflann_impl.hpp:1335-1347, 987-1143).
//
// The structure is an octree over Morton prefixes (see common.cuh).  It is built without any
// level-by-level dependency and without a host round trip, from the sorted codes alone:
//
//   1. k_bounds    cloud bounding box (block reduction + ordered-int atomics) + finite check
//   2. k_morton    30-bit Morton key of every point; thread 0 publishes the lattice {lo, scale}
//   3. radix sort  (key, original index) pairs (prims.cu; steps 1-3 are ONE kernel of a 16-CTA cluster for
//                  scan-sized clouds, cluster_sort.cu)
//   4. k_cells     per point: the level L(i) of its leaf cell = 1 + the longest prefix that any
//                  window of kLeafMax+1 consecutive sorted points containing i still shares; the gather of the point
//                  into Morton order; per block and level t = 0..9 the number of points that are the first point of
//                  an internal cell of level t.  The LAST block to finish turns these counts into per-block
//                  prefixes and per-level bases: the level-t node that holds point i gets the breadth-first id
//                  level_base[t] + block_prefix[t][block of i] + (cells of level t begun in i's block up to i) - 1,
//                  and the node count stays on the device
//   5. k_init_nodes  empty boxes (+inf / -inf) and references for the nodes that exist (for scan-sized clouds:
//                  for the whole node array, on a side stream, while the sort runs on its 16 SMs)
//   6. k_emit      for every level, every point folds its coordinates into the slot of
//                  its cell in the parent node.  Points are sorted, so the points of one cell are
//                  consecutive lanes: a segmented warp reduction first, then ONE float atomic
//                  per run and box component.  The first point of a cell writes the child reference
//                  (leaf: start, the count is accumulated per run; internal: child id + meta).
//
// The node array is sized by the bound "every internal cell holds more than kLeafMax points and the
// cells of one level are disjoint": at most 10 n / 17 + 1 nodes; only the nodes that exist are touched.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "prims.cuh"

namespace ddlo {

int morton_sort_cluster(ddlo_runtime* rt, const float4* pts, int n, unsigned* keys_out, int* vals_out, float* lattice);  // cluster_sort.cu
constexpr int kClusterSortMax = 16 * 512 * 16;  // the largest cloud the cluster sort takes

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u ^ 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
  return __uint_as_float(u);
}

// bounds: [0..2] ordered min, [3..5] ordered max, [6] non-finite counter
__global__ void __launch_bounds__(256) k_bounds(const float4* __restrict__ pts, int n, unsigned* __restrict__ bounds) {
  __shared__ float s_lo[8][3], s_hi[8][3];
  __shared__ unsigned s_bad[8];
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  unsigned bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
      bad = 1;
      continue;
    }
    lo[0] = fminf(lo[0], p.x);
    lo[1] = fminf(lo[1], p.y);
    lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x);
    hi[1] = fmaxf(hi[1], p.y);
    hi[2] = fmaxf(hi[2], p.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) {
      s_lo[warp][a] = lo[a];
      s_hi[warp][a] = hi[a];
    }
    s_bad[warp] = bad;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float l = s_lo[0][threadIdx.x], h = s_hi[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) {
      l = fminf(l, s_lo[w][threadIdx.x]);
      h = fmaxf(h, s_hi[w][threadIdx.x]);
    }
    atomicMin(bounds + threadIdx.x, f2ord(l));
    atomicMax(bounds + 3 + threadIdx.x, f2ord(h));
  }
  if (threadIdx.x == 3) {
    unsigned b = 0;
    for (int w = 0; w < 8; ++w) b |= s_bad[w];
    if (b) atomicAdd(bounds + 6, 1u);
  }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__device__ __forceinline__ unsigned compact10(unsigned v) {  // inverse of spread10
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030c30c3u;
  v = (v | (v >> 4)) & 0x0300f00fu;
  v = (v | (v >> 8)) & 0x030000ffu;
  v = (v | (v >> 16)) & 0x3ffu;
  return v;
}

// lattice: {lo.x, lo.y, lo.z, scale, (int) non-finite count, (int) node count [written later]}
__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ pts, int n, const unsigned* __restrict__ bounds,
                                                unsigned* __restrict__ keys, int* __restrict__ vals, float* __restrict__ lattice) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lx = ord2f(bounds[0]), ly = ord2f(bounds[1]), lz = ord2f(bounds[2]);
  const float ex = ord2f(bounds[3]) - lx, ey = ord2f(bounds[4]) - ly, ez = ord2f(bounds[5]) - lz;
  const float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-30f));
  const float scale = 1023.0f / ext;  // one isotropic lattice: cells stay cubes
  if (i == 0) {
    lattice[0] = lx;
    lattice[1] = ly;
    lattice[2] = lz;
    lattice[3] = scale;
    reinterpret_cast<unsigned*>(lattice)[4] = bounds[6];
  }
  const float4 p = pts[i];
  const unsigned cx = (unsigned)fminf(fmaxf((p.x - lx) * scale, 0.0f), 1023.0f);
  const unsigned cy = (unsigned)fminf(fmaxf((p.y - ly) * scale, 0.0f), 1023.0f);
  const unsigned cz = (unsigned)fminf(fmaxf((p.z - lz) * scale, 0.0f), 1023.0f);
  keys[i] = spread10(cx) | (spread10(cy) << 1) | (spread10(cz) << 2);
  vals[i] = i;
}

// number of leading 3-bit digits two 30-bit codes share (10 if equal)
__device__ __forceinline__ int common_digits(unsigned a, unsigned b) {
  const unsigned x = a ^ b;
  return x ? (__clz(x) - 2) / 3 : kMortonLevels;
}

// float min / max by integer atomics: non-negative floats order like ints, negative ones in reverse like unsigned ints.
// Correct for any mix of signs when the cell starts at +inf (min) / -inf (max); -0 is folded into +0 first.
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  v += 0.0f;
  if (v >= 0.0f)
    atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  v += 0.0f;
  if (v >= 0.0f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// per-build control words (device): [0] ticket of k_cells' blocks, [1] node count, [2..11] level bases

// inclusive count, over the threads of the block up to and including this one, of a per-thread flag; all threads call
__device__ __forceinline__ int block_inclusive_count(bool flag, int* s_wtot /*[8]*/, int* block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) s_wtot[warp] = __popc(b);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int t = s_wtot[w];
    if (w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  if (block_total) *block_total = tot;
  return base + __popc(b & (0xffffffffu >> (31 - lane)));
}

__global__ void __launch_bounds__(256) k_cells(const unsigned* __restrict__ keys, const int* __restrict__ perm,
                                               const float4* __restrict__ pts, int n, unsigned char* __restrict__ leaf_level,
                                               int* __restrict__ counts /*[10][gridDim.x]: in: -, out: per-block prefixes*/, int* __restrict__ ctl,
                                               float4* __restrict__ spts, float* __restrict__ lattice) {
  __shared__ int s_wtot[8];
  __shared__ int s_last;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  int L = 0, cd = kMortonLevels;
  if (live) {
    const unsigned key = keys[i];
    // longest prefix shared by kLeafMax+1 consecutive points around i: that cell is too big for a leaf
    int g = -1;
    const int j0 = max(0, i - kLeafMax), j1 = min(i, n - 1 - kLeafMax);
    for (int j = j0; j <= j1; ++j) g = max(g, common_digits(__ldg(keys + j), __ldg(keys + j + kLeafMax)));
    L = min(kMortonLevels, max(1, g + 1));
    leaf_level[i] = (unsigned char)L;
    cd = i == 0 ? -1 : common_digits(key, __ldg(keys + i - 1));
    const int o = perm[i];
    DDLO_CHECK_INDEX(o, n, "k_cells: permutation entry");
    const float4 p = pts[o];
    spts[i] = make_float4(p.x, p.y, p.z, __int_as_float(o));
  }
  // cells of level t that begin in this block: one ballot per level and warp, the warp totals added in shared memory
  __shared__ int s_cnt[kMortonLevels];
  if (threadIdx.x < kMortonLevels) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < kMortonLevels; ++t) {
      const unsigned b = __ballot_sync(0xffffffffu, live && t < L && cd < t);
      if (lane == 0 && b) atomicAdd(&s_cnt[t], __popc(b));
    }
  }
  __syncthreads();
  if (threadIdx.x < kMortonLevels) counts[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x];
  // the last block to arrive turns the counts into exclusive prefixes over the blocks and publishes the level bases
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ctl, 1) == (int)gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // one warp per level (levels warp, warp + 8): a running exclusive scan over the blocks, 32 at a time, no block barrier
  const int nb = (int)gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < kMortonLevels; t += 8) {
    int* row = counts + (size_t)t * nb;
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 256) {  // a lane takes 8 consecutive blocks: 8 independent loads in flight
      const int first = b0 + lane * 8;
      int c[8], sum = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        c[k] = first + k < nb ? __ldcg(row + first + k) : 0;
        sum += c[k];
      }
      int inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      int run = carry + inc - sum;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (first + k < nb) row[first + k] = run;
        run += c[k];
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_cnt[t] = carry;  // cells of the level
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int level_base = 0;
    for (int t = 0; t < kMortonLevels; ++t) {
      ctl[2 + t] = level_base;
      level_base += s_cnt[t];
    }
    ctl[1] = level_base;  // node count
    reinterpret_cast<int*>(lattice)[5] = level_base;
    ctl[0] = 0;
  }
}

// empty nodes: boxes lo = +inf, hi = -inf, references 0.  n_nodes_ptr == nullptr: the whole array of max_nodes nodes.
__global__ void __launch_bounds__(256) k_init_nodes(unsigned* __restrict__ words, const int* __restrict__ n_nodes_ptr, size_t max_nodes) {
  const size_t n_words = (n_nodes_ptr ? (size_t)(*n_nodes_ptr) : max_nodes) * 64;
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n_words; w += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(w & 63);
    words[w] = k < 24 ? 0x7f800000u : (k < 48 ? 0xff800000u : 0u);
  }
}

// One thread per sorted point.  For every tree level t = 1..L(i) the point belongs to a cell that is
// a child slot of the level-(t-1) node holding it.  The points of that cell are consecutive, so
// inside a warp they form one run of lanes: the run's box (and, for a leaf, its size) is reduced
// with shuffles and its first lane issues the atomics.
__global__ void __launch_bounds__(256) k_emit(const unsigned* __restrict__ keys, const unsigned char* __restrict__ leaf_level,
                                              const int* __restrict__ block_prefix /*[10][gridDim.x]*/, const int* __restrict__ ctl,
                                              const float4* __restrict__ spts, int n, float* __restrict__ nodes, int4* __restrict__ meta,
                                              int* __restrict__ node_of_point) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = i < n;
  if (i == 0) meta[0] = make_int4(-1, 0, 0, 0);  // root
  unsigned key = 0;
  int L = 0, cd = kMortonLevels;
  float lo[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  float hi[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
  int orig = 0;
  if (live) {
    key = keys[i];
    L = leaf_level[i];
    cd = i == 0 ? -1 : common_digits(key, __ldg(keys + i - 1));
    const float4 p = spts[i];
    orig = __float_as_int(p.w);
    lo[0] = hi[0] = p.x;
    lo[1] = hi[1] = p.y;
    lo[2] = hi[2] = p.z;
  }
  // deepest leaf level inside the block (all warps walk the same number of levels: the loop holds block barriers)
  __shared__ int s_lmax;
  __shared__ int s_cell0;
  __shared__ float s_box[8][6];
  __shared__ int s_wtot[8];
  if (threadIdx.x == 0) s_lmax = 0;
  __syncthreads();
  {
    const int wl = __reduce_max_sync(0xffffffffu, L);
    if (lane == 0) atomicMax(&s_lmax, wl);
  }
  __syncthreads();
  const int Lmax = s_lmax;
  // The node ids of the point's cells on all its levels: breadth-first = level base + cells of the level begun in
  // earlier blocks + cells begun in this block up to the point (k_cells left the first two on the device).
  int nid[kMortonLevels];
#pragma unroll
  for (int t = 0; t < kMortonLevels; ++t) {
    nid[t] = -1;
    if (t >= Lmax) continue;  // uniform: nobody in the block has a cell on this level
    const int within = block_inclusive_count(live && t < L && cd < t, s_wtot, nullptr);
    if (live && t < L) nid[t] = __ldg(ctl + 2 + t) + __ldg(block_prefix + (size_t)t * gridDim.x + blockIdx.x) + within - 1;
  }
#pragma unroll
  for (int t = 1; t <= kMortonLevels; ++t) {
    if (t > Lmax) break;
    const bool in = live && t <= L;
    const int parent = in ? nid[t - 1] : -1;
    const int slot = (key >> (3 * (kMortonLevels - t))) & 7;
    const int cell = in ? parent * 8 + slot : -1 - lane;  // run id; dead lanes never join a run
    if (in) DDLO_CHECK_INDEX(parent, __ldg(ctl + 1), "k_emit: parent node");
    if (in && t == L) {
      DDLO_CHECK_INDEX(orig, n, "k_emit: original index");
      node_of_point[orig] = parent;
    }
    // child reference, written by the first point of the cell
    if (in && cd < t) {
      unsigned* pw = reinterpret_cast<unsigned*>(nodes) + (size_t)parent * 64;
      if (t < L) {
        const int child = nid[t < kMortonLevels ? t : 0];  // t < L <= 10 here
        pw[48 + 2 * slot] = (unsigned)child;
        pw[49 + 2 * slot] = 0xffffffffu;
        const unsigned keep = ~((1u << (kMortonLevels - t)) - 1u) & 0x3ffu;  // the t leading bits of each axis
        const unsigned origin = (compact10(key) & keep) | ((compact10(key >> 1) & keep) << 10) | ((compact10(key >> 2) & keep) << 20);
        meta[child] = make_int4(parent, t, (int)origin, 0);
      } else {
        pw[48 + 2 * slot] = (unsigned)i;  // leaf start; the count is accumulated below
      }
    }
    // High in the tree all 256 points of a block sit in ONE internal cell: then the block folds its box with one
    // atomic per component instead of one per warp (at the root's children 2048 warps would queue on 48 words).
    if (threadIdx.x == 0) s_cell0 = cell;
    __syncthreads();
    const bool whole_block = __syncthreads_and(in && t < L && cell == s_cell0) != 0;
    if (whole_block) {
      const int warp = threadIdx.x >> 5;
      float r[6];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        r[a] = ord2f(__reduce_min_sync(0xffffffffu, f2ord(lo[a])));
        r[3 + a] = ord2f(__reduce_max_sync(0xffffffffu, f2ord(hi[a])));
      }
      if (lane < 6) s_box[warp][lane] = r[lane];
      __syncthreads();
      if (threadIdx.x < 6) {
        float acc = s_box[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) acc = threadIdx.x < 3 ? fminf(acc, s_box[w][threadIdx.x]) : fmaxf(acc, s_box[w][threadIdx.x]);
        float* pw = nodes + (size_t)parent * 64 + 8 * threadIdx.x + slot;
        if (threadIdx.x < 3)
          atomic_min_float(pw, acc);
        else
          atomic_max_float(pw, acc);
      }
      continue;
    }
    // segmented reduction over runs of equal `cell` (runs are contiguous lanes)
    float v[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
    int cnt = 1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int oc = __shfl_down_sync(0xffffffffu, cell, d);
      const int on = __shfl_down_sync(0xffffffffu, cnt, d);
      float ov[6];
#pragma unroll
      for (int a = 0; a < 6; ++a) ov[a] = __shfl_down_sync(0xffffffffu, v[a], d);
      if (lane + d < 32 && oc == cell) {
        cnt += on;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          v[a] = fminf(v[a], ov[a]);
          v[3 + a] = fmaxf(v[3 + a], ov[3 + a]);
        }
      }
    }
    const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
    if (in && (lane == 0 || prev != cell)) {
      float* pw = nodes + (size_t)parent * 64;
      atomic_min_float(pw + slot, v[0]);
      atomic_min_float(pw + 8 + slot, v[1]);
      atomic_min_float(pw + 16 + slot, v[2]);
      atomic_max_float(pw + 24 + slot, v[3]);
      atomic_max_float(pw + 32 + slot, v[4]);
      atomic_max_float(pw + 40 + slot, v[5]);
      if (t == L) atomicAdd(reinterpret_cast<unsigned*>(pw) + 49 + 2 * slot, (unsigned)cnt);
    }
  }
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

static int build_index_impl(ddlo_cloud* c, char** base_out) {
  ddlo_runtime* rt = c->rt;
  const int n = c->n;
  cudaStream_t st = rt->stream;
  const int tb = 256;
  const int nb = (n + tb - 1) / tb;

  // temporaries: keys | keys_alt | vals | vals_alt | leaf_level | counts[10][nb] | sort scratch, from the stream-ordered
  // pool (warmed at runtime creation): no cudaMalloc in a frame
  const size_t sz = align256((size_t)n * 4);
  const size_t off_lvl = 4 * sz, off_counts = off_lvl + align256((size_t)n);
  const size_t off_sort = off_counts + align256((size_t)kMortonLevels * nb * sizeof(int));
  char* base = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&base), off_sort + radix_sort_temp_bytes(n) + 256, st));
  *base_out = base;
  unsigned* keys = reinterpret_cast<unsigned*>(base);
  unsigned* keys_alt = reinterpret_cast<unsigned*>(base + sz);
  int* vals = reinterpret_cast<int*>(base + 2 * sz);
  int* vals_alt = reinterpret_cast<int*>(base + 3 * sz);
  unsigned char* leaf_level = reinterpret_cast<unsigned char*>(base + off_lvl);
  int* counts = reinterpret_cast<int*>(base + off_counts);
  char* sort_tmp = base + off_sort;
  unsigned* bounds = reinterpret_cast<unsigned*>(sort_tmp + radix_sort_temp_bytes(n));  // 8 words, generic path only
  // control words of this build (ticket, node count, level bases): a zeroed corner of the runtime's scratch; the
  // last block of k_cells leaves the ticket at zero again
  int* ctl = reinterpret_cast<int*>(static_cast<char*>(rt->d_scratch) + 4096);

  const size_t max_nodes = (size_t)kMortonLevels * n / (kLeafMax + 1) + 1;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->spts), (size_t)n * sizeof(float4), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->nodes), max_nodes * 256, st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->meta), max_nodes * sizeof(int4), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->node_of_point), (size_t)n * sizeof(int), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->lattice), 8 * sizeof(float), st));
  unsigned* words = reinterpret_cast<unsigned*>(c->nodes);
  const int gb = std::min(std::max(nb, 1), rt->num_sms * 4);

  // Scan-sized clouds: box, Morton keys and the sort in one kernel of a 16-CTA cluster (cluster_sort.cu), and the
  // node array is initialised meanwhile on the runtime's side stream (the cluster occupies 16 of the SMs).
  // Larger clouds, or a device that refuses the cluster launch: the generic kernels + the radix sort of prims.cu.
  static const bool cluster_sort_enabled = std::getenv("DDLO_NO_CLUSTER_SORT") == nullptr;
  const unsigned* skeys = keys;
  const int* svals = vals;
  bool nodes_ready = false;
  if (cluster_sort_enabled && n <= kClusterSortMax && rt->side) {
    DDLO_CUDA(cudaEventRecord(rt->ev_fork, st));  // behind the allocation of the node array
    DDLO_CUDA(cudaStreamWaitEvent(rt->side, rt->ev_fork, 0));
    k_init_nodes<<<gb, tb, 0, rt->side>>>(words, nullptr, max_nodes);
    DDLO_CUDA(cudaEventRecord(rt->ev_join, rt->side));
    rt->launches += 1;
    nodes_ready = true;
  }
  if (!(cluster_sort_enabled && morton_sort_cluster(rt, c->pts, n, keys, vals, c->lattice) == DDLO_OK)) {
    DDLO_CUDA(cudaMemsetAsync(bounds, 0xff, 12, st));
    DDLO_CUDA(cudaMemsetAsync(bounds + 3, 0x00, 16, st));
    k_bounds<<<std::min(nb, rt->num_sms), tb, 0, st>>>(c->pts, n, bounds);
    k_morton<<<nb, tb, 0, st>>>(c->pts, n, bounds, keys, vals, c->lattice);
    rt->launches += 2;
    unsigned* ks = nullptr;
    int* vs = nullptr;
    DDLO_TRY(radix_sort_pairs(st, keys, keys_alt, vals, vals_alt, n, 30, sort_tmp, &ks, &vs, &rt->launches));
    skeys = ks;
    svals = vs;
  }
  k_cells<<<nb, tb, 0, st>>>(skeys, svals, c->pts, n, leaf_level, counts, ctl, c->spts, c->lattice);
  if (nodes_ready) {
    DDLO_CUDA(cudaStreamWaitEvent(st, rt->ev_join, 0));
  } else {
    k_init_nodes<<<gb, tb, 0, st>>>(words, ctl + 1, 0);
    rt->launches += 1;
  }
  k_emit<<<nb, tb, 0, st>>>(skeys, leaf_level, counts, ctl, c->spts, n, reinterpret_cast<float*>(c->nodes), c->meta, c->node_of_point);
  rt->launches += 2;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

int build_index(ddlo_cloud* c) {
  if (c->has_index) return DDLO_OK;
  ddlo_runtime* rt = c->rt;
  const int n = c->n;
  if (n <= 0) return fail(DDLO_E_EMPTY, "build_index: empty cloud");
  if (n > (1 << 26)) return fail(DDLO_E_UNSUPPORTED, "build_index: cloud too large");
  if (c->shared) return fail(DDLO_E_INVALID, "build_index: the cloud is shared (immutable) and has no index");
  char* base = nullptr;
  const int rc = build_index_impl(c, &base);
  if (base) cudaFreeAsync(base, rt->stream);
  if (rc != DDLO_OK) {  // nothing half-built stays behind
    const std::string why = ddlo_last_error();
    for (void** p : {reinterpret_cast<void**>(&c->spts), reinterpret_cast<void**>(&c->nodes), reinterpret_cast<void**>(&c->meta),
                     reinterpret_cast<void**>(&c->node_of_point), reinterpret_cast<void**>(&c->lattice)}) {
      if (*p) cudaFreeAsync(*p, rt->stream);
      *p = nullptr;
    }
    return fail(rc, why);
  }
  c->view.spts = c->spts;
  c->view.nodes = c->nodes;
  c->view.meta = c->meta;
  c->view.node_of_point = c->node_of_point;
  c->view.lattice = c->lattice;
  c->view.n = n;
  c->has_index = true;
  return DDLO_OK;
}

// non-finite coordinate count of an indexed cloud (one small read-back; used by the checked entry points)
int index_nonfinite_count(ddlo_cloud* c, int* count) {
  *count = 0;
  if (!c->has_index) return DDLO_OK;
  DDLO_CUDA(cudaMemcpyAsync(c->rt->h_pinned, c->lattice, 8 * sizeof(float), cudaMemcpyDeviceToHost, c->rt->stream));
  DDLO_CUDA(cudaStreamSynchronize(c->rt->stream));
  *count = static_cast<int*>(c->rt->h_pinned)[4];
  return DDLO_OK;
}

}  // namespace ddlo
