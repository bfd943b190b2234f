// kNN index construction on the device: replaces the serial kd-tree build of
// KdTreeFLANN::setInputCloud -> KDTreeSingleIndexAdaptor::buildIndex / divideTree
// (R/nanoflann.hpp:137-143, R/impl/nanoflann_impl.hpp:1335-1347, 987-1143).
//
// The structure is an octree over Morton prefixes (see common.cuh).  It is built without any
// level-by-level dependency and without a host round trip, from the sorted codes alone:
//
//   1. k_bounds    cloud bounding box (block reduction + ordered-int atomics) + finite check
//   2. k_morton    30-bit Morton key of every point; thread 0 publishes the lattice {lo, scale}
//   3. radix sort  (key, original index) pairs                         [cub::DeviceRadixSort]
//   4. k_cells     per point: the level L(i) of its leaf cell = 1 + the longest prefix that any
//                  window of kLeafMax+1 consecutive sorted points containing i still shares; the
//                  flag "i is the first point of an internal cell of level t" for t = 0..9; the
//                  gather of the point into Morton order
//   5. inclusive scan of the level-major flags                         [cub::DeviceScan]
//                  -> breadth-first node ids: the level-t node holding point i is S[t][i] - 1;
//                  the last element is the node count (stays on the device)
//   6. k_init_nodes, k_emit: for every level, every point folds its coordinates into the slot of
//                  its cell in the parent node.  Points are sorted, so the points of one cell are
//                  consecutive lanes: a segmented warp reduction first, then ONE ordered-int atomic
//                  per run and box component.  The first point of a cell writes the child reference
//                  (leaf: start, the count is accumulated per run; internal: child id + meta).
//   7. k_finalize  ordered ints -> floats
//
// The node array is sized by the bound "every internal cell holds more than kLeafMax points and the
// cells of one level are disjoint": at most 10 n / 17 + 1 nodes; only the nodes that exist are touched.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace ddlo {

int morton_sort_cluster(ddlo_runtime* rt, const float4* pts, int n, unsigned* keys_out, int* vals_out, float* lattice);  // cluster_sort.cu

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u ^ 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
  return __uint_as_float(u);
}
constexpr unsigned kOrdPosInf = 0xff800000u;  // f2ord(+inf)
constexpr unsigned kOrdNegInf = 0x007fffffu;  // f2ord(-inf)

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

__global__ void __launch_bounds__(256) k_cells(const unsigned* __restrict__ keys, const int* __restrict__ perm,
                                               const float4* __restrict__ pts, int n, unsigned char* __restrict__ leaf_level,
                                               int* __restrict__ flags /*[10][n]*/, float4* __restrict__ spts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned key = keys[i];
  // longest prefix shared by kLeafMax+1 consecutive points around i: that cell is too big for a leaf
  int g = -1;
  const int j0 = max(0, i - kLeafMax), j1 = min(i, n - 1 - kLeafMax);
  for (int j = j0; j <= j1; ++j) g = max(g, common_digits(__ldg(keys + j), __ldg(keys + j + kLeafMax)));
  const int L = min(kMortonLevels, max(1, g + 1));
  leaf_level[i] = (unsigned char)L;
  const int cd = i == 0 ? -1 : common_digits(key, __ldg(keys + i - 1));
#pragma unroll
  for (int t = 0; t < kMortonLevels; ++t) flags[(size_t)t * n + i] = (t < L && cd < t) ? 1 : 0;
  const int o = perm[i];
  const float4 p = pts[o];
  spts[i] = make_float4(p.x, p.y, p.z, __int_as_float(o));
}

// the node count is the last element of the scanned flags
__global__ void __launch_bounds__(256) k_init_nodes(unsigned* __restrict__ words, const int* __restrict__ n_nodes_ptr, float* lattice) {
  const size_t n_words = (size_t)(*n_nodes_ptr) * 64;
  if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<int*>(lattice)[5] = *n_nodes_ptr;
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n_words; w += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(w & 63);
    words[w] = k < 24 ? kOrdPosInf : (k < 48 ? kOrdNegInf : 0u);
  }
}

// One thread per sorted point.  For every tree level t = 1..L(i) the point belongs to a cell that is
// a child slot of the level-(t-1) node S[t-1][i]-1.  The points of that cell are consecutive, so
// inside a warp they form one run of lanes: the run's box (and, for a leaf, its size) is reduced
// with shuffles and its first lane issues the atomics.
__global__ void __launch_bounds__(256) k_emit(const unsigned* __restrict__ keys, const unsigned char* __restrict__ leaf_level,
                                              const int* __restrict__ S /*[10][n] inclusive scan*/, const float4* __restrict__ spts, int n,
                                              unsigned* __restrict__ nodes, int4* __restrict__ meta, int* __restrict__ node_of_point) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = i < n;
  if (i == 0) meta[0] = make_int4(-1, 0, 0, 0);  // root
  unsigned key = 0;
  int L = 0, cd = kMortonLevels;
  unsigned o[6] = {kOrdPosInf, kOrdPosInf, kOrdPosInf, kOrdNegInf, kOrdNegInf, kOrdNegInf};
  int orig = 0;
  if (live) {
    key = keys[i];
    L = leaf_level[i];
    cd = i == 0 ? -1 : common_digits(key, __ldg(keys + i - 1));
    const float4 p = spts[i];
    orig = __float_as_int(p.w);
    o[0] = o[3] = f2ord(p.x);
    o[1] = o[4] = f2ord(p.y);
    o[2] = o[5] = f2ord(p.z);
  }
  // deepest leaf level inside the block (all warps walk the same number of levels: the loop holds block barriers)
  __shared__ int s_lmax;
  __shared__ int s_cell0;
  __shared__ unsigned s_box[8][6];
  if (threadIdx.x == 0) s_lmax = 0;
  __syncthreads();
  {
    const int wl = __reduce_max_sync(0xffffffffu, L);
    if (lane == 0) atomicMax(&s_lmax, wl);
  }
  __syncthreads();
  const int Lmax = s_lmax;
  // the node ids of the point's cells on all its levels, fetched up front (independent loads: one memory
  // latency instead of one per level)
  int nid[kMortonLevels];
#pragma unroll
  for (int t = 0; t < kMortonLevels; ++t) nid[t] = (live && t < L) ? __ldg(S + (size_t)t * n + i) - 1 : -1;
#pragma unroll
  for (int t = 1; t <= kMortonLevels; ++t) {
    if (t > Lmax) break;
    const bool in = live && t <= L;
    const int parent = in ? nid[t - 1] : -1;
    const int slot = (key >> (3 * (kMortonLevels - t))) & 7;
    const int cell = in ? parent * 8 + slot : -1 - lane;  // run id; dead lanes never join a run
    if (in && t == L) node_of_point[orig] = parent;
    // child reference, written by the first point of the cell
    if (in && cd < t) {
      unsigned* pw = nodes + (size_t)parent * 64;
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
      unsigned r[6];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        r[a] = __reduce_min_sync(0xffffffffu, o[a]);
        r[3 + a] = __reduce_max_sync(0xffffffffu, o[3 + a]);
      }
      if (lane < 6) s_box[warp][lane] = r[lane];
      __syncthreads();
      if (threadIdx.x < 6) {
        unsigned acc = s_box[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) acc = threadIdx.x < 3 ? min(acc, s_box[w][threadIdx.x]) : max(acc, s_box[w][threadIdx.x]);
        unsigned* pw = nodes + (size_t)parent * 64 + 8 * threadIdx.x + slot;
        if (threadIdx.x < 3)
          atomicMin(pw, acc);
        else
          atomicMax(pw, acc);
      }
      continue;
    }
    // segmented reduction over runs of equal `cell` (runs are contiguous lanes)
    unsigned v[6] = {o[0], o[1], o[2], o[3], o[4], o[5]};
    int cnt = 1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int oc = __shfl_down_sync(0xffffffffu, cell, d);
      const int on = __shfl_down_sync(0xffffffffu, cnt, d);
      unsigned ov[6];
#pragma unroll
      for (int a = 0; a < 6; ++a) ov[a] = __shfl_down_sync(0xffffffffu, v[a], d);
      if (lane + d < 32 && oc == cell) {
        cnt += on;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          v[a] = min(v[a], ov[a]);
          v[3 + a] = max(v[3 + a], ov[3 + a]);
        }
      }
    }
    const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
    if (in && (lane == 0 || prev != cell)) {
      unsigned* pw = nodes + (size_t)parent * 64;
      atomicMin(pw + slot, v[0]);
      atomicMin(pw + 8 + slot, v[1]);
      atomicMin(pw + 16 + slot, v[2]);
      atomicMax(pw + 24 + slot, v[3]);
      atomicMax(pw + 32 + slot, v[4]);
      atomicMax(pw + 40 + slot, v[5]);
      if (t == L) atomicAdd(pw + 49 + 2 * slot, (unsigned)cnt);
    }
  }
}

__global__ void __launch_bounds__(256) k_finalize(unsigned* __restrict__ words, const int* __restrict__ n_nodes_ptr) {
  const size_t n_words = (size_t)(*n_nodes_ptr) * 64;
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < n_words; w += (size_t)gridDim.x * blockDim.x)
    if ((w & 63) < 48) words[w] = __float_as_uint(ord2f(words[w]));
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int build_index(ddlo_cloud* c) {
  if (c->has_index) return DDLO_OK;
  ddlo_runtime* rt = c->rt;
  const int n = c->n;
  if (n <= 0) return fail(DDLO_E_EMPTY, "build_index: empty cloud");
  if (n > (1 << 26)) return fail(DDLO_E_UNSUPPORTED, "build_index: cloud too large");
  cudaStream_t st = rt->stream;

  // temporaries: bounds (8 u32) | keys | keys_alt | vals | vals_alt | leaf_level | flags[10][n] | cub temp
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DoubleBuffer<unsigned> kb0(nullptr, nullptr);
  cub::DoubleBuffer<int> vb0(nullptr, nullptr);
  DDLO_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, kb0, vb0, n, 0, 30, st));
  const size_t n_flags = (size_t)kMortonLevels * n;
  DDLO_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, (long long)n_flags, st));
  const size_t cub_bytes = std::max(sort_bytes, scan_bytes);
  const size_t off_keys = 256, sz = align256((size_t)n * 4);
  const size_t off_lvl = off_keys + 4 * sz, off_flags = off_lvl + align256((size_t)n);
  const size_t off_cub = off_flags + align256(n_flags * 4);
  // temporaries come from the stream-ordered pool (warmed at runtime creation): no cudaMalloc in a frame
  char* base = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&base), off_cub + align256(cub_bytes), st));
  unsigned* bounds = reinterpret_cast<unsigned*>(base);
  unsigned* keys = reinterpret_cast<unsigned*>(base + off_keys);
  unsigned* keys_alt = reinterpret_cast<unsigned*>(base + off_keys + sz);
  int* vals = reinterpret_cast<int*>(base + off_keys + 2 * sz);
  int* vals_alt = reinterpret_cast<int*>(base + off_keys + 3 * sz);
  unsigned char* leaf_level = reinterpret_cast<unsigned char*>(base + off_lvl);
  int* flags = reinterpret_cast<int*>(base + off_flags);
  void* cub_tmp = base + off_cub;

  const size_t max_nodes = (size_t)kMortonLevels * n / (kLeafMax + 1) + 1;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->spts), (size_t)n * sizeof(float4), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->nodes), max_nodes * 256, st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->meta), max_nodes * sizeof(int4), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->node_of_point), (size_t)n * sizeof(int), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->lattice), 8 * sizeof(float), st));

  const int tb = 256;
  const int nb = (n + tb - 1) / tb;
  // Scan-sized clouds: box, Morton keys and the sort in one kernel of a 16-CTA cluster (cluster_sort.cu).
  // Larger clouds, or a device that refuses the cluster launch: the generic kernels + CUB's radix sort.
  static const bool cluster_sort_enabled = std::getenv("DDLO_NO_CLUSTER_SORT") == nullptr;
  const unsigned* skeys = keys;
  const int* svals = vals;
  if (!(cluster_sort_enabled && morton_sort_cluster(rt, c->pts, n, keys, vals, c->lattice) == DDLO_OK)) {
    DDLO_CUDA(cudaMemsetAsync(bounds, 0xff, 12, st));
    DDLO_CUDA(cudaMemsetAsync(bounds + 3, 0x00, 16, st));
    k_bounds<<<std::min(nb, rt->num_sms), tb, 0, st>>>(c->pts, n, bounds);
    k_morton<<<nb, tb, 0, st>>>(c->pts, n, bounds, keys, vals, c->lattice);
    cub::DoubleBuffer<unsigned> kb(keys, keys_alt);
    cub::DoubleBuffer<int> vb(vals, vals_alt);
    DDLO_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, sort_bytes, kb, vb, n, 0, 30, st));
    // (the radix-sort and scan passes are CUB library kernels and are not counted as ours)
    skeys = kb.Current();
    svals = vb.Current();
    rt->launches += 2;
  }
  k_cells<<<nb, tb, 0, st>>>(skeys, svals, c->pts, n, leaf_level, flags, c->spts);
  DDLO_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, scan_bytes, flags, flags, (long long)n_flags, st));
  const int* n_nodes_ptr = flags + (n_flags - 1);
  unsigned* words = reinterpret_cast<unsigned*>(c->nodes);
  const int gb = std::min(nb, rt->num_sms * 4);
  k_init_nodes<<<gb, tb, 0, st>>>(words, n_nodes_ptr, c->lattice);
  k_emit<<<nb, tb, 0, st>>>(skeys, leaf_level, flags, c->spts, n, words, c->meta, c->node_of_point);
  k_finalize<<<gb, tb, 0, st>>>(words, n_nodes_ptr);
  rt->launches += 4;
  DDLO_CUDA(cudaFreeAsync(base, st));
  DDLO_CUDA(cudaGetLastError());

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
