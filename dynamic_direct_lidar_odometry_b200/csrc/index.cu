// kNN index construction on the device: replaces the serial kd-tree build of
// KdTreeFLANN::setInputCloud -> KDTreeSingleIndexAdaptor::buildIndex / divideTree
// (R/nanoflann.hpp:137-143, R/impl/nanoflann_impl.hpp:1335-1347, 987-1143).
//
//   1. k_bounds   cloud bounding box (ordered-int atomics) + finite check
//   2. k_morton   30-bit Morton key of every point on a uniform 1024^3 lattice over the box
//   3. radix sort (key, original index) pairs                         [cub::DeviceRadixSort]
//   4. k_leaves   gather points into Morton order (w := original index) and box every 8 of them
//   5. k_level    box every 8 boxes, once per upper level
#include <cub/device/device_radix_sort.cuh>

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

__device__ __forceinline__ void store_box(float4* __restrict__ level, int node, float lx, float ly, float lz, float hx, float hy,
                                          float hz) {
  float* g = reinterpret_cast<float*>(level + (size_t)(node >> 3) * 12) + (node & 7);
  g[0] = lx;
  g[8] = ly;
  g[16] = lz;
  g[24] = hx;
  g[32] = hy;
  g[40] = hz;
}

// one thread per padded point slot (npad = groups * 64); 8 consecutive lanes form one leaf
__global__ void __launch_bounds__(256) k_leaves(const float4* __restrict__ pts, const int* __restrict__ perm, int n, int npad,
                                                float4* __restrict__ spts, float4* __restrict__ leaf_level) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;  // npad is a multiple of 64 and blockDim of 32: whole warps leave together
  const float inf = __int_as_float(0x7f800000);
  float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
  float4 out = make_float4(inf, inf, inf, __int_as_float(-1));
  if (i < n) {
    const int o = perm[i];
    const float4 p = pts[o];
    out = make_float4(p.x, p.y, p.z, __int_as_float(o));
    lx = hx = p.x;
    ly = hy = p.y;
    lz = hz = p.z;
  }
  spts[i] = out;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o));
    ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
    lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
    hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
    hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
    hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
  }
  if ((i & 7) == 0) store_box(leaf_level, i >> 3, lx, ly, lz, hx, hy, hz);
}

// one thread per node slot of `level` (slots = groups * 8); children are group `node` of `child_level`
__global__ void __launch_bounds__(256) k_level(const float4* __restrict__ child_level, int child_groups, float4* __restrict__ level,
                                               int slots) {
  const int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= slots) return;
  const float inf = __int_as_float(0x7f800000);
  float lx = inf, ly = inf, lz = inf, hx = -inf, hy = -inf, hz = -inf;
  if (node < child_groups) {
    const float4* g = child_level + (size_t)node * 12;
    const float4 a0 = g[0], a1 = g[1], b0 = g[2], b1 = g[3], c0 = g[4], c1 = g[5];
    const float4 d0 = g[6], d1 = g[7], e0 = g[8], e1 = g[9], f0 = g[10], f1 = g[11];
    lx = fminf(fminf(fminf(a0.x, a0.y), fminf(a0.z, a0.w)), fminf(fminf(a1.x, a1.y), fminf(a1.z, a1.w)));
    ly = fminf(fminf(fminf(b0.x, b0.y), fminf(b0.z, b0.w)), fminf(fminf(b1.x, b1.y), fminf(b1.z, b1.w)));
    lz = fminf(fminf(fminf(c0.x, c0.y), fminf(c0.z, c0.w)), fminf(fminf(c1.x, c1.y), fminf(c1.z, c1.w)));
    hx = fmaxf(fmaxf(fmaxf(d0.x, d0.y), fmaxf(d0.z, d0.w)), fmaxf(fmaxf(d1.x, d1.y), fmaxf(d1.z, d1.w)));
    hy = fmaxf(fmaxf(fmaxf(e0.x, e0.y), fmaxf(e0.z, e0.w)), fmaxf(fmaxf(e1.x, e1.y), fmaxf(e1.z, e1.w)));
    hz = fmaxf(fmaxf(fmaxf(f0.x, f0.y), fmaxf(f0.z, f0.w)), fmaxf(fmaxf(f1.x, f1.y), fmaxf(f1.z, f1.w)));
  }
  store_box(level, node, lx, ly, lz, hx, hy, hz);
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
  cudaStream_t st = rt->stream;

  // level sizes: cnt[L-1] = leaves, cnt[l] = ceil(cnt[l+1] / 8), cnt[0] <= 8
  int cnt_rev[kMaxLevels];
  int nlev = 0;
  int m = (n + kLeaf - 1) / kLeaf;
  for (;;) {
    if (nlev >= kMaxLevels) return fail(DDLO_E_UNSUPPORTED, "build_index: cloud too large");
    cnt_rev[nlev++] = m;
    if (m <= kBranch) break;
    m = (m + kBranch - 1) / kBranch;
  }
  int cnt[kMaxLevels], groups[kMaxLevels];
  size_t total_groups = 0;
  for (int l = 0; l < nlev; ++l) {
    cnt[l] = cnt_rev[nlev - 1 - l];
    groups[l] = (cnt[l] + kBranch - 1) / kBranch;
    total_groups += groups[l];
  }
  const int npad = groups[nlev - 1] * kBranch * kLeaf;

  // scratch: bounds (8 u32) | keys | keys_alt | vals | vals_alt | cub temp
  size_t cub_bytes = 0;
  cub::DoubleBuffer<unsigned> kb0(nullptr, nullptr);
  cub::DoubleBuffer<int> vb0(nullptr, nullptr);
  DDLO_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, kb0, vb0, n, 0, 30, st));
  const size_t off_keys = 256, sz = align256((size_t)n * 4);
  const size_t need = off_keys + 4 * sz + align256(cub_bytes);
  DDLO_TRY(ensure_scratch(rt, need));
  char* base = static_cast<char*>(rt->d_scratch);
  unsigned* bounds = reinterpret_cast<unsigned*>(base);
  unsigned* keys = reinterpret_cast<unsigned*>(base + off_keys);
  unsigned* keys_alt = reinterpret_cast<unsigned*>(base + off_keys + sz);
  int* vals = reinterpret_cast<int*>(base + off_keys + 2 * sz);
  int* vals_alt = reinterpret_cast<int*>(base + off_keys + 3 * sz);
  void* cub_tmp = base + off_keys + 4 * sz;

  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->spts), (size_t)npad * sizeof(float4), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&c->boxes), total_groups * 12 * sizeof(float4), st));

  DDLO_CUDA(cudaMemsetAsync(bounds, 0xff, 12, st));
  DDLO_CUDA(cudaMemsetAsync(bounds + 3, 0x00, 16, st));
  const int tb = 256;
  const int gb = std::min((n + tb - 1) / tb, rt->num_sms * 8);
  k_bounds<<<gb, tb, 0, st>>>(c->pts, n, bounds);
  k_morton<<<(n + tb - 1) / tb, tb, 0, st>>>(c->pts, n, bounds, keys, vals);
  rt->launches += 2;
  cub::DoubleBuffer<unsigned> kb(keys, keys_alt);
  cub::DoubleBuffer<int> vb(vals, vals_alt);
  DDLO_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kb, vb, n, 0, 30, st));
  rt->launches += 4;  // onesweep: histogram + 4 digit passes are library kernels; counted loosely

  float4* lvl_ptr[kMaxLevels];
  {
    float4* p = c->boxes;
    for (int l = 0; l < nlev; ++l) {
      lvl_ptr[l] = p;
      p += (size_t)groups[l] * 12;
    }
  }
  k_leaves<<<(npad + tb - 1) / tb, tb, 0, st>>>(c->pts, vb.Current(), n, npad, c->spts, lvl_ptr[nlev - 1]);
  rt->launches += 1;
  for (int l = nlev - 2; l >= 0; --l) {
    const int slots = groups[l] * kBranch;
    k_level<<<(slots + tb - 1) / tb, tb, 0, st>>>(lvl_ptr[l + 1], groups[l + 1], lvl_ptr[l], slots);
    rt->launches += 1;
  }
  DDLO_CUDA(cudaGetLastError());

  c->npad = npad;
  c->view.spts = c->spts;
  c->view.n = n;
  c->view.nlev = nlev;
  for (int l = 0; l < kMaxLevels; ++l) {
    c->view.box[l] = l < nlev ? lvl_ptr[l] : nullptr;
    c->view.cnt[l] = l < nlev ? cnt[l] : 0;
  }
  c->has_index = true;
  return DDLO_OK;
}

}  // namespace ddlo
