// Scan preprocessing on the device: the two PCL filters OdomNode::preprocessPoints applies right
// before the registration path (odom.cc:442-478, configured at odom.cc:115-130), and that it applies
// to every new keyframe (odom.cc:494-499, 1133-1137).  SURVEY.md §8f row 2.  At the end of the file: the
// residual image OdomNode builds right after the registration (odom.cc:804-827, §8f row 3).
//
//   voxel_filter   pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.10, filters/impl/voxel_grid.hpp; PCL is
//                  not vendored by the reference, its published algorithm is restated): bounding box of
//                  the finite points, voxel index idx = (floor(p * inv_leaf) - min_b) . (1, dx, dx*dy),
//                  points sorted by idx, one output point per occupied voxel = the float centroid
//                  sum / count, voxels in ascending idx order.
//   crop_box       pcl::CropBox<PointXYZI>::applyFilter (filters/impl/crop_box.hpp) with an identity
//                  box pose: a point is inside iff min <= p <= max on every axis; `negative` keeps the
//                  outside; `keep_organized` replaces removed points by NaN instead of compacting.
//
// PCL sorts the (idx, point) pairs with std::sort, so the order in which the points of a voxel are
// added (and with it the last bits of the float centroid) is unspecified there.  Here the sort is
// stable: the points of a voxel are added in their original order, one thread per voxel, which makes
// the result deterministic and bit-identical to the CPU oracle's restatement.
#include "prims.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "common.cuh"

namespace ddlo {

__device__ __forceinline__ unsigned pp_f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u ^ 0x80000000u);
}
static inline float pp_ord2f_host(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// box[0..2] ordered min, box[3..5] ordered max, box[6] number of finite points
__global__ void __launch_bounds__(256) k_pp_bounds(const float4* __restrict__ pts, int n, unsigned* __restrict__ box) {
  __shared__ unsigned s[8][7];
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u}, cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
    const unsigned o[3] = {pp_f2ord(p.x), pp_f2ord(p.y), pp_f2ord(p.z)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = min(lo[a], o[a]);
      hi[a] = max(hi[a], o[a]);
    }
    cnt += 1;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) {
      s[warp][a] = lo[a];
      s[warp][3 + a] = hi[a];
    }
    s[warp][6] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    unsigned v = s[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? min(v, s[w][threadIdx.x]) : (threadIdx.x < 6 ? max(v, s[w][threadIdx.x]) : v + s[w][threadIdx.x]);
    if (threadIdx.x < 3)
      atomicMin(box + threadIdx.x, v);
    else if (threadIdx.x < 6)
      atomicMax(box + threadIdx.x, v);
    else
      atomicAdd(box + 6, v);
  }
}

struct VoxelGridSpec {
  float inv[3];      // inverse_leaf_size_
  int min_b[3];      // min_b_
  int mul[3];        // divb_mul_
  unsigned invalid;  // key of non-finite points: the number of voxels of the grid, sorts behind every voxel
};

// voxel_grid.hpp: ijk = static_cast<int>(std::floor(p * inverse_leaf_size) - static_cast<float>(min_b)); idx = ijk . divb_mul
__global__ void __launch_bounds__(256) k_voxel_keys(const float4* __restrict__ pts, int n, VoxelGridSpec g, unsigned* __restrict__ keys,
                                                     int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  unsigned key = g.invalid;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    const int i0 = (int)(floorf(__fmul_rn(p.x, g.inv[0])) - (float)g.min_b[0]);
    const int i1 = (int)(floorf(__fmul_rn(p.y, g.inv[1])) - (float)g.min_b[1]);
    const int i2 = (int)(floorf(__fmul_rn(p.z, g.inv[2])) - (float)g.min_b[2]);
    key = (unsigned)(i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);
  }
  keys[i] = key;
  vals[i] = i;
}

__global__ void __launch_bounds__(256) k_voxel_heads(const unsigned* __restrict__ keys, int n, unsigned invalid, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned k = keys[i];
  flags[i] = (k != invalid && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

// one thread per occupied voxel: float sums in the (stable) sorted order, then sum / count
__global__ void __launch_bounds__(256) k_voxel_centroids(const unsigned* __restrict__ keys, const int* __restrict__ perm,
                                                          const int* __restrict__ slot /* inclusive scan of the head flags */,
                                                          const float4* __restrict__ pts, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned k = keys[i];
  const int s = slot[i];
  if (s == 0 || (i > 0 && slot[i - 1] == s)) return;  // not the first point of a voxel
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  int cnt = 0;
  for (int j = i; j < n && keys[j] == k; ++j) {
    const float4 p = pts[perm[j]];
    sx = __fadd_rn(sx, p.x);
    sy = __fadd_rn(sy, p.y);
    sz = __fadd_rn(sz, p.z);
    ++cnt;
  }
  const float c = (float)cnt;
  out[s - 1] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), 1.0f);
}

static inline size_t pp_align256(size_t x) { return (x + 255) & ~size_t(255); }

// *d_out (cudaMallocAsync on the runtime's stream, the caller owns it) and *n_out receive the filtered cloud
int voxel_filter_device(ddlo_runtime* rt, const float4* pts, int n, const float leaf[3], float4** d_out, int* n_out) {
  *d_out = nullptr;
  *n_out = 0;
  if (!(leaf[0] > 0.0f && leaf[1] > 0.0f && leaf[2] > 0.0f)) return fail(DDLO_E_INVALID, "voxel filter: leaf size must be positive");
  cudaStream_t st = rt->stream;
  const int tb = 256, nb = std::max(1, (n + tb - 1) / tb);
  const size_t sort_bytes = radix_sort_temp_bytes(std::max(n, 1)), scan_bytes = scan_temp_bytes((size_t)std::max(n, 1));
  const size_t sz = pp_align256((size_t)std::max(n, 1) * 4);
  const size_t total = 256 + 5 * sz + pp_align256(std::max(sort_bytes, scan_bytes));
  char* base = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&base), total, st));
  unsigned* box = reinterpret_cast<unsigned*>(base);
  unsigned* keys = reinterpret_cast<unsigned*>(base + 256);
  unsigned* keys_alt = reinterpret_cast<unsigned*>(base + 256 + sz);
  int* vals = reinterpret_cast<int*>(base + 256 + 2 * sz);
  int* vals_alt = reinterpret_cast<int*>(base + 256 + 3 * sz);
  int* flags = reinterpret_cast<int*>(base + 256 + 4 * sz);
  void* cub_tmp = base + 256 + 5 * sz;
  auto cleanup = [&]() { cudaFreeAsync(base, st); };

  // bounding box of the finite points (getMinMax3D)
  unsigned h_box[7];
  DDLO_CUDA(cudaMemsetAsync(box, 0xff, 12, st));
  DDLO_CUDA(cudaMemsetAsync(box + 3, 0x00, 16, st));
  if (n > 0) {
    k_pp_bounds<<<std::min(nb, rt->num_sms * 2), tb, 0, st>>>(pts, n, box);
    rt->launches += 1;
  }
  DDLO_CUDA(cudaMemcpyAsync(h_box, box, sizeof(h_box), cudaMemcpyDeviceToHost, st));
  DDLO_CUDA(cudaStreamSynchronize(st));
  if (h_box[6] == 0) {  // nothing finite: empty output
    cleanup();
    DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(d_out), sizeof(float4), st));
    return DDLO_OK;
  }
  VoxelGridSpec g;
  long long div[3];
  for (int a = 0; a < 3; ++a) {
    g.inv[a] = 1.0f / leaf[a];
    const float lo = pp_ord2f_host(h_box[a]), hi = pp_ord2f_host(h_box[3 + a]);
    // the check PCL makes before anything else: the index space must fit an int
    g.min_b[a] = (int)std::floor(lo * g.inv[a]);
    const int max_b = (int)std::floor(hi * g.inv[a]);
    div[a] = (long long)max_b - (long long)g.min_b[a] + 1;
  }
  if (div[0] * div[1] * div[2] > (long long)std::numeric_limits<int>::max()) {
    cleanup();
    return fail(DDLO_E_UNSUPPORTED, "voxel filter: leaf size is too small for the input dataset, integer indices would overflow");
  }
  g.mul[0] = 1;
  g.mul[1] = (int)div[0];
  g.mul[2] = (int)(div[0] * div[1]);
  g.invalid = (unsigned)(div[0] * div[1] * div[2]);
  int bits = 1;
  while (bits < 32 && (1ull << bits) <= (unsigned long long)g.invalid) ++bits;

  k_voxel_keys<<<nb, tb, 0, st>>>(pts, n, g, keys, vals);
  unsigned* skeys = nullptr;
  int* svals = nullptr;
  {
    const int rc = radix_sort_pairs(st, keys, keys_alt, vals, vals_alt, n, bits, cub_tmp, &skeys, &svals, &rt->launches);  // stable: original order inside a voxel
    if (rc != DDLO_OK) {
      cleanup();
      return rc;
    }
  }
  k_voxel_heads<<<nb, tb, 0, st>>>(skeys, n, g.invalid, flags);
  {
    const int rc = scan_int(st, flags, flags, (size_t)n, true, cub_tmp, &rt->launches);
    if (rc != DDLO_OK) {
      cleanup();
      return rc;
    }
  }
  int n_vox = 0;
  DDLO_CUDA(cudaMemcpyAsync(&n_vox, flags + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  DDLO_CUDA(cudaStreamSynchronize(st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(d_out), std::max<size_t>(1, (size_t)n_vox) * sizeof(float4), st));
  k_voxel_centroids<<<nb, tb, 0, st>>>(skeys, svals, flags, pts, n, *d_out);
  rt->launches += 3;
  cleanup();
  DDLO_CUDA(cudaGetLastError());
  *n_out = n_vox;
  return DDLO_OK;
}

// ---- crop box ---------------------------------------------------------------------------------------
struct CropSpec {
  float lo[3], hi[3];
  int negative;
};
__device__ __forceinline__ bool crop_keeps(const float4 p, const CropSpec& c) {
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return false;  // non-finite points are dropped (is_dense == false path)
  const bool outside = (p.x < c.lo[0] || p.y < c.lo[1] || p.z < c.lo[2]) || (p.x > c.hi[0] || p.y > c.hi[1] || p.z > c.hi[2]);
  return c.negative ? outside : !outside;
}
__global__ void __launch_bounds__(256) k_crop_flags(const float4* __restrict__ pts, int n, CropSpec c, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = crop_keeps(pts[i], c) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_crop_compact(const float4* __restrict__ pts, int n, const int* __restrict__ slot, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = slot[i];
  if (s != (i == 0 ? 0 : slot[i - 1])) {
    const float4 p = pts[i];
    out[s - 1] = make_float4(p.x, p.y, p.z, 1.0f);
  }
}
__global__ void __launch_bounds__(256) k_crop_organized(const float4* __restrict__ pts, int n, CropSpec c, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const float nan = __int_as_float(0x7fc00000);
  out[i] = crop_keeps(p, c) ? make_float4(p.x, p.y, p.z, 1.0f) : make_float4(nan, nan, nan, 1.0f);  // user_filter_value_ = NaN
}

int crop_box_device(ddlo_runtime* rt, const float4* pts, int n, const float lo[3], const float hi[3], int negative, int keep_organized,
                    float4** d_out, int* n_out) {
  *d_out = nullptr;
  *n_out = 0;
  cudaStream_t st = rt->stream;
  CropSpec c;
  for (int a = 0; a < 3; ++a) {
    c.lo[a] = lo[a];
    c.hi[a] = hi[a];
  }
  c.negative = negative ? 1 : 0;
  const int tb = 256, nb = std::max(1, (n + tb - 1) / tb);
  if (keep_organized || n == 0) {
    DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(d_out), std::max<size_t>(1, (size_t)n) * sizeof(float4), st));
    if (n > 0) {
      k_crop_organized<<<nb, tb, 0, st>>>(pts, n, c, *d_out);
      rt->launches += 1;
    }
    DDLO_CUDA(cudaGetLastError());
    *n_out = n;
    return DDLO_OK;
  }
  const size_t scan_bytes = scan_temp_bytes((size_t)n);
  const size_t sz = pp_align256((size_t)n * 4);
  char* base = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&base), sz + pp_align256(scan_bytes), st));
  int* flags = reinterpret_cast<int*>(base);
  k_crop_flags<<<nb, tb, 0, st>>>(pts, n, c, flags);
  DDLO_TRY(scan_int(st, flags, flags, (size_t)n, true, base + sz, &rt->launches));
  int kept = 0;
  DDLO_CUDA(cudaMemcpyAsync(&kept, flags + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  DDLO_CUDA(cudaStreamSynchronize(st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(d_out), std::max<size_t>(1, (size_t)kept) * sizeof(float4), st));
  k_crop_compact<<<nb, tb, 0, st>>>(pts, n, flags, *d_out);
  rt->launches += 2;
  cudaFreeAsync(base, st);
  DDLO_CUDA(cudaGetLastError());
  *n_out = kept;
  return DDLO_OK;
}

// ---- strided organised down-sample (first stage) ---------------------------------
// OdomNode::preprocessPoints, odom.cc:445-455: pcl::ExtractIndices<PointType> with the index mask built at
// odom.cc:124-130 (every downsample_filter_row_-th row and downsample_filter_col_-th column of the
// cloud_height_ x cloud_width_ scan), setNegative(false), setKeepOrganized(true): the cloud keeps its size and
// order, and every point that is NOT in the mask gets user_filter_value_ = NaN in all its fields
// (pcl/filters/impl/extract_indices.hpp, PCL 1.10: `output = *input_`, then the removed indices are overwritten).
__global__ void __launch_bounds__(256) k_extract_stride(const float4* __restrict__ pts, int n, int width, int height, int row_stride,
                                                        int col_stride, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int row = i / width, col = i - row * width;
  const bool keep = row < height && row % row_stride == 0 && col % col_stride == 0;
  const float nan = __int_as_float(0x7fc00000);
  const float4 p = pts[i];
  out[i] = keep ? make_float4(p.x, p.y, p.z, 1.0f) : make_float4(nan, nan, nan, 1.0f);
}

int extract_stride_device(ddlo_runtime* rt, const float4* pts, int n, int width, int height, int row_stride, int col_stride, float4** d_out) {
  *d_out = nullptr;
  cudaStream_t st = rt->stream;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(d_out), std::max<size_t>(1, (size_t)n) * sizeof(float4), st));
  if (n > 0) {
    k_extract_stride<<<(n + 255) / 256, 256, 0, st>>>(pts, n, width, height, row_stride, col_stride, *d_out);
    rt->launches += 1;
  }
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

// ---- residual image ------------------------------------------------------------
// OdomNode::scanMatching, odom.cc:804-827: the registration scan is projected into a W x H angular image,
// theta = atan2(x, z), phi = atan2(y, sqrt(x^2 + z^2)), u = int((theta - a_min) / (a_max - a_min) * W), v alike
// with H (C++ int conversion: truncation toward zero), and cell (v, u) keeps x, y, z and the GICP residual
// (as intensity) of the LAST scan point that falls into it; the other cells stay zero.  The reference's
// sequential loop makes "last" the largest point index; here an atomicMax per cell elects it.  The angles are
// FLOAT: `atan2(pt.x, pt.z)` and `sqrt(pt.x * pt.x + pt.z * pt.z)` have float arguments, and odom.cc sees
// `using namespace std;` (odom.h -> detection/detection.h -> tracking/tracking.h -> tracking/hungarian.h:42), so the
// float overloads are chosen and only their results are widened to double.  atan2 is evaluated in double and
// rounded to float here, i.e. the correctly rounded float value.
struct ResidualImageSpec {
  int w, h;
  double a_min, a_span;
};
__device__ __forceinline__ int residual_cell(const float4 p, const ResidualImageSpec& g) {
  const double theta = (double)(float)atan2((double)p.x, (double)p.z);
  const float rxz = __fsqrt_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.z, p.z)));
  const double phi = (double)(float)atan2((double)p.y, (double)rxz);
  const int u = (int)((theta - g.a_min) / g.a_span * (double)g.w);
  const int v = (int)((phi - g.a_min) / g.a_span * (double)g.h);
  if (u < 0 || u >= g.w || v < 0 || v >= g.h) return -1;
  return v * g.w + u;
}
__global__ void __launch_bounds__(256) k_residual_owner(const float4* __restrict__ pts, int n, ResidualImageSpec g, int* __restrict__ owner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return;  // atan2 of NaN gives NaN, whose int conversion is undefined: skipped
  const int c = residual_cell(p, g);
  if (c >= 0) atomicMax(owner + c, i);
}
__global__ void __launch_bounds__(256) k_residual_fill(const float4* __restrict__ pts, const float* __restrict__ sqd, int cells,
                                                        const int* __restrict__ owner, float4* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const int i = owner[c];
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i >= 0) {
    const float4 p = pts[i];
    o = make_float4(p.x, p.y, p.z, (float)sqrt((double)sqd[i]));  // residuals[i] = sqrt(sq_distances_[i]) (double), stored as float intensity
  }
  out[c] = o;
}

// d_out: w*h float4 (x, y, z, residual) on the device
int residual_image_device(ddlo_runtime* rt, const float4* pts, const float* sqd, int n, int w, int h, double a_min, double a_max, float4* d_out) {
  cudaStream_t st = rt->stream;
  const int cells = w * h;
  int* owner = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&owner), (size_t)cells * sizeof(int), st));
  DDLO_CUDA(cudaMemsetAsync(owner, 0xff, (size_t)cells * sizeof(int), st));  // -1
  ResidualImageSpec g{w, h, a_min, a_max - a_min};
  if (n > 0) k_residual_owner<<<(n + 255) / 256, 256, 0, st>>>(pts, n, g, owner);
  k_residual_fill<<<(cells + 255) / 256, 256, 0, st>>>(pts, sqd, cells, owner, d_out);
  rt->launches += 2;
  cudaFreeAsync(owner, st);
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

}  // namespace ddlo
