// kNN index construction on the device: replaces the serial kd-tree build of
// KdTreeFLANN::setInputCloud -> KDTreeSingleIndexAdaptor::buildIndex / divideTree
// (R/nanoflann.hpp:137-143, R/impl/nanoflann_impl.hpp:1335-1347, 987-1143).
//
// The structure is an octree over Morton prefixes (see common.cuh).  It is built without any
// level-by-level dependency, from the sorted codes alone:
//
//   1. k_bounds    cloud bounding box (ordered-int atomics) + finite check
//   2. k_morton    30-bit Morton key of every point
//   3. radix sort  (key, original index) pairs                         [cub::DeviceRadixSort]
//   4. k_cells     per point: the level L(i) of its leaf cell = 1 + the longest prefix that any
//                  window of kLeafMax+1 consecutive sorted points containing i still shares; the
//                  flag "i is the first point of an internal cell of level t" for t = 0..9; the
//                  gather of the point into Morton order
//   5. inclusive scan of the level-major flags                         [cub::DeviceScan]
//                  -> breadth-first node ids: the level-t node holding point i is S[t][i] - 1
//   6. k_init_nodes, k_emit: every first point of a cell writes that cell into its parent node
//                  (leaf: range + box; internal: child id) and every leaf folds its box into the
//                  slots of all its ancestors with ordered-int atomics
//   7. k_finalize  ordered ints -> floats
//
// One 4-byte read-back (the node count, after step 5) sizes the node array exactly.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstring>

#include "common.cuh"

namespace ddlo {

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
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(bounds + a, f2ord(lo[a]));
      atomicMax(bounds + 3 + a, f2ord(hi[a]));
    }
    if (bad) atomicAdd(bounds + 6, 1u);
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

__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ pts, int n, const unsigned* __restrict__ bounds,
                                                unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lx = ord2f(bounds[0]), ly = ord2f(bounds[1]), lz = ord2f(bounds[2]);
  const float ex = ord2f(bounds[3]) - lx, ey = ord2f(bounds[4]) - ly, ez = ord2f(bounds[5]) - lz;
  const float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-30f));
  const float scale = 1023.0f / ext;  // one isotropic lattice: cells stay cubes
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

// one thread per 32-bit word of the node array
__global__ void __launch_bounds__(256) k_init_nodes(unsigned* __restrict__ words, size_t n_words) {
  const size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const int k = (int)(w & 63);
  words[w] = k < 24 ? kOrdPosInf : (k < 48 ? kOrdNegInf : 0u);
}

__device__ __forceinline__ void merge_box(unsigned* node_words, int slot, const unsigned o[6]) {
  atomicMin(node_words + slot, o[0]);
  atomicMin(node_words + 8 + slot, o[1]);
  atomicMin(node_words + 16 + slot, o[2]);
  atomicMax(node_words + 24 + slot, o[3]);
  atomicMax(node_words + 32 + slot, o[4]);
  atomicMax(node_words + 40 + slot, o[5]);
}

__device__ __forceinline__ unsigned compact10(unsigned v) {  // inverse of spread10
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030c30c3u;
  v = (v | (v >> 4)) & 0x0300f00fu;
  v = (v | (v >> 8)) & 0x030000ffu;
  v = (v | (v >> 16)) & 0x3ffu;
  return v;
}

__global__ void __launch_bounds__(256) k_emit(const unsigned* __restrict__ keys, const unsigned char* __restrict__ leaf_level,
                                              const int* __restrict__ S /*[10][n] inclusive scan*/, const float4* __restrict__ spts, int n,
                                              unsigned* __restrict__ nodes, int4* __restrict__ meta, int* __restrict__ node_of_point) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) meta[0] = make_int4(-1, 0, 0, 0);  // root
  const unsigned key = keys[i];
  const int L = leaf_level[i];
  const int cd = i == 0 ? -1 : common_digits(key, __ldg(keys + i - 1));
  // i is the first point of its level-t cell for every t > cd; cells of level t <= L exist in the tree
  for (int t = max(1, cd + 1); t <= L; ++t) {
    const int parent = S[(size_t)(t - 1) * n + i] - 1;
    const int slot = (key >> (3 * (kMortonLevels - t))) & 7;
    unsigned* pw = nodes + (size_t)parent * 64;
    if (t < L) {  // internal child: its box is assembled by the leaves below it
      const int child = S[(size_t)t * n + i] - 1;
      pw[48 + 2 * slot] = (unsigned)child;
      pw[49 + 2 * slot] = 0xffffffffu;
      const unsigned keep = ~((1u << (kMortonLevels - t)) - 1u) & 0x3ffu;  // the t leading bits of each axis
      const unsigned origin = (compact10(key) & keep) | ((compact10(key >> 1) & keep) << 10) | ((compact10(key >> 2) & keep) << 20);
      meta[child] = make_int4(parent, t, (int)origin, 0);
      continue;
    }
    // leaf: the run of points sharing the first t digits
    int e = i + 1;
    while (e < n && common_digits(__ldg(keys + e), key) >= t) ++e;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int j = i; j < e; ++j) {
      const float4 p = spts[j];
      node_of_point[__float_as_int(p.w)] = parent;
      lo[0] = fminf(lo[0], p.x), lo[1] = fminf(lo[1], p.y), lo[2] = fminf(lo[2], p.z);
      hi[0] = fmaxf(hi[0], p.x), hi[1] = fmaxf(hi[1], p.y), hi[2] = fmaxf(hi[2], p.z);
    }
    const unsigned o[6] = {f2ord(lo[0]), f2ord(lo[1]), f2ord(lo[2]), f2ord(hi[0]), f2ord(hi[1]), f2ord(hi[2])};
    pw[48 + 2 * slot] = (unsigned)i;
    pw[49 + 2 * slot] = (unsigned)(e - i);
    merge_box(pw, slot, o);
    for (int u = t - 1; u >= 1; --u) {
      const int anc_parent = S[(size_t)(u - 1) * n + i] - 1;
      merge_box(nodes + (size_t)anc_parent * 64, (key >> (3 * (kMortonLevels - u))) & 7, o);
    }
  }
}

__global__ void __launch_bounds__(256) k_finalize(unsigned* __restrict__ words, size_t n_words) {
  const size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  if ((w & 63) < 48) words[w] = __float_as_uint(ord2f(words[w]));
}

static float host_ord2f(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
  float f;
  std::memcpy(&f, &u, sizeof(f));
  return f;
}

static int ensure_scratch(ddlo_runtime* rt, size_t bytes) {
  if (rt->d_scratch_bytes >= bytes) return DDLO_OK;
  if (rt->d_scratch) {
    DDLO_CUDA(cudaStreamSynchronize(rt->stream));
    DDLO_CUDA(cudaFree(rt->d_scratch));
    rt->d_scratch = nullptr;
    rt->d_scratch_bytes = 0;
  }
  const size_t want = bytes + bytes / 4 + (1u << 20);
  DDLO_CUDA(cudaMalloc(&rt->d_scratch, want));
  rt->d_scratch_bytes = want;
  return DDLO_OK;
}

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int build_index(ddlo_cloud* c) {
  if (c->has_index) return DDLO_OK;
  ddlo_runtime* rt = c->rt;
  const int n = c->n;
  if (n <= 0) return fail(DDLO_E_EMPTY, "build_index: empty cloud");
  if (n > (1 << 30)) return fail(DDLO_E_UNSUPPORTED, "build_index: cloud too large");
  cudaStream_t st = rt->stream;

  // scratch: bounds (8 u32) | keys | keys_alt | vals | vals_alt | leaf_level | flags[10][n] | cub temp
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
  DDLO_TRY(ensure_scratch(rt, off_cub + align256(cub_bytes)));
  char* base = static_cast<char*>(rt->d_scratch);
  unsigned* bounds = reinterpret_cast<unsigned*>(base);
  unsigned* keys = reinterpret_cast<unsigned*>(base + off_keys);
  unsigned* keys_alt = reinterpret_cast<unsigned*>(base + off_keys + sz);
  int* vals = reinterpret_cast<int*>(base + off_keys + 2 * sz);
  int* vals_alt = reinterpret_cast<int*>(base + off_keys + 3 * sz);
  unsigned char* leaf_level = reinterpret_cast<unsigned char*>(base + off_lvl);
  int* flags = reinterpret_cast<int*>(base + off_flags);
  void* cub_tmp = base + off_cub;

  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->spts), (size_t)n * sizeof(float4), st));

  DDLO_CUDA(cudaMemsetAsync(bounds, 0xff, 12, st));
  DDLO_CUDA(cudaMemsetAsync(bounds + 3, 0x00, 16, st));
  const int tb = 256;
  const int nb = (n + tb - 1) / tb;
  k_bounds<<<std::min(nb, rt->num_sms * 8), tb, 0, st>>>(c->pts, n, bounds);
  k_morton<<<nb, tb, 0, st>>>(c->pts, n, bounds, keys, vals);
  cub::DoubleBuffer<unsigned> kb(keys, keys_alt);
  cub::DoubleBuffer<int> vb(vals, vals_alt);
  DDLO_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, sort_bytes, kb, vb, n, 0, 30, st));
  // (the radix-sort and scan passes are CUB library kernels and are not counted as ours)
  k_cells<<<nb, tb, 0, st>>>(kb.Current(), vb.Current(), c->pts, n, leaf_level, flags, c->spts);
  DDLO_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, scan_bytes, flags, flags, (long long)n_flags, st));
  rt->launches += 3;
  DDLO_CUDA(cudaGetLastError());

  // the one read-back of the build: how many nodes the tree has, and the lattice of the codes
  int* h_count = static_cast<int*>(rt->h_pinned);
  unsigned* h_bounds = reinterpret_cast<unsigned*>(h_count + 4);
  DDLO_CUDA(cudaMemcpyAsync(h_count, flags + (n_flags - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  DDLO_CUDA(cudaMemcpyAsync(h_bounds, bounds, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  DDLO_CUDA(cudaStreamSynchronize(st));
  const int n_nodes = *h_count;
  if (n_nodes < 1) return fail(DDLO_E_CUDA, "build_index: node count came back empty");
  if (h_bounds[6] != 0u) return fail(DDLO_E_NONFINITE, "build_index: the cloud holds NaN or Inf coordinates");
  float blo[3], bhi[3];
  for (int a = 0; a < 3; ++a) {
    blo[a] = host_ord2f(h_bounds[a]);
    bhi[a] = host_ord2f(h_bounds[3 + a]);
  }
  // same float expressions as k_morton
  const float ext = std::fmax(std::fmax(bhi[0] - blo[0], bhi[1] - blo[1]), std::fmax(bhi[2] - blo[2], 1e-30f));
  const float scale = 1023.0f / ext;

  const size_t n_words = (size_t)n_nodes * 64;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->nodes), n_words * 4, st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->meta), (size_t)n_nodes * sizeof(int4), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->node_of_point), (size_t)n * sizeof(int), st));
  unsigned* words = reinterpret_cast<unsigned*>(c->nodes);
  const int wb = (int)((n_words + tb - 1) / tb);
  k_init_nodes<<<wb, tb, 0, st>>>(words, n_words);
  k_emit<<<nb, tb, 0, st>>>(kb.Current(), leaf_level, flags, c->spts, n, words, c->meta, c->node_of_point);
  k_finalize<<<wb, tb, 0, st>>>(words, n_words);
  rt->launches += 3;
  DDLO_CUDA(cudaGetLastError());

  c->view.spts = c->spts;
  c->view.nodes = c->nodes;
  c->view.meta = c->meta;
  c->view.node_of_point = c->node_of_point;
  for (int a = 0; a < 3; ++a) c->view.lo[a] = blo[a];
  c->view.scale = scale;
  c->view.n = n;
  c->view.n_nodes = n_nodes;
  c->has_index = true;
  return DDLO_OK;
}

}  // namespace ddlo
