// Range-image segmentation on the device: the per-frame stage of the reference's
// DetectionModule that follows the registration -- projectScan (src/detection/detection.cpp:254-329),
// groundRemoval (:448-512), cloudSegmentation (:514-546) and labelComponents (:548-724).
//
// The reference labels segments with one sequential queue flood fill per seed, seeds in raster order.  Three of its
// outputs depend on the order in which that queue pushes pixels: the min_z / max_z pair (a new minimum never counts
// for the maximum, :612-616), the float sum of the residuals (:632-634) and, through both, which segments are
// accepted and how they are numbered.  The order is therefore reproduced exactly, but not sequentially:
//
//   k_seg_project   range, ground flag and the initial label of every pixel (pixel-parallel; the column-wise
//                   bottom-up loop of groundRemoval collapses to a function of two vertical pixel pairs)
//   k_seg_edges     the flood fill's admission test for the 4 neighbours of every pixel -> 4 bits per pixel.
//                   The test is symmetric, so segments are the connected components of an undirected graph:
//   k_ccl_*         union-find labelling (horizontal runs are joined by a ballot inside each warp first); the root of a
//                   component is its smallest raster index = the seed the reference starts it from.  Component sizes
//                   give every segment its slice of the push list.
//   k_seg_fill      one warp per segment replays the queue: 8 queue entries x 4 neighbours per step, a neighbour
//                   wanted twice in a step goes to the earlier (entry, direction) -- which is what the queue order
//                   means -- and the survivors are appended in lane order.  The edge bits of the whole image and the
//                   visited bits sit in shared memory, the last 1024 queue entries in a per-warp ring, so a step never
//                   waits on global memory.  The same warp then walks its push list once more for the statistics:
//                   prefix-minimum for the min_z rule, a lane-ordered add chain for the residual sum.
//   k_seg_labels    accepted segments numbered by an exclusive scan over the seeds in raster order.
//
// Arithmetic: float products/sums rounded one by one as the reference's -O2 x86-64 build does; atan2 is evaluated in
// double and rounded to float (the correctly rounded float value but for ~1e-9 of the inputs; the reference's libm
// atan2f may differ from it in the last bit, which matters only for a slope within 1 ulp of its threshold).
#include <vector>

#include "prims.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace ddlo {

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kFillThreads = 256;
constexpr int kFillWarps = kFillThreads / 32;
constexpr int kRing = 1024;  // queue entries a warp keeps in shared memory

struct SegDev {
  int H, W, HW;
  int ground_rows;
  int wr0, wr1, wc0, wc1;
  int valid_point_num, min_line_num, valid_line_num;
  float minimum_range, ground_thr, mount, theta;
  float sin_x, cos_x, sin_y, cos_y;
  float x0, y0, z0, height;
  float min_delta_z, max_delta_z, max_distance, max_elevation;
  int have_residuals;
  int unordered_sums;
};

__device__ __forceinline__ bool seg_in_window(const SegDev& p, int y, int x) { return y >= p.wr0 && y <= p.wr1 && x >= p.wc0 && x <= p.wc1; }

// full_cloud_ / range_mat_ of one pixel (:305-326): false = the pixel keeps the NaN point and range 0
__device__ __forceinline__ bool seg_full_point(const float* __restrict__ scan, int stride, const SegDev& p, int idx, float& x, float& y, float& z,
                                               float& range) {
  const float* q = scan + (size_t)idx * stride;
  x = q[0], y = q[1], z = q[2];
  range = 0.0f;
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) return false;
  const float dx = __fadd_rn(x, p.x0), dy = __fadd_rn(y, p.y0), dz = __fadd_rn(z, p.z0);
  const float r = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  if (r < p.minimum_range) return false;
  range = r;
  return true;
}

// one iteration of the groundRemoval loop for the pair (row, row - 1) of a column (:470-493):
// -1 "no info", 1 ground (marks both rows), 0 neither
__device__ int seg_pair_state(const float* __restrict__ scan, int stride, const SegDev& p, int row, int col) {
  float lx, ly, lz, lr, ux, uy, uz, ur;
  const bool vl = seg_full_point(scan, stride, p, row * p.W + col, lx, ly, lz, lr);
  const bool vu = seg_full_point(scan, stride, p, (row - 1) * p.W + col, ux, uy, uz, ur);
  if ((vl && lx == 0.0f) || (vu && ux == 0.0f)) return -1;  // an invalid pixel holds NaN, and NaN == 0 is false
  if (!vl || !vu) return 0;                                 // NaN slope: the comparison below is false
  const float dx = __fsub_rn(ux, lx), dy = __fsub_rn(uy, ly), dz = __fsub_rn(uz, lz);
  const float h = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  const float a = (float)atan2((double)dz, (double)h);
  const float angle = (float)((double)__fmul_rn(a, 180.0f) / 3.14159265358979323846);
  return fabsf(__fsub_rn(angle, p.mount)) <= p.ground_thr ? 1 : 0;
}

// OdomNode::transformScans for the segmentation scan (odom.cc:957-963): pcl::transformPointCloud in float,
// (r0 x + r1 y) + (r2 z + t) per row; points with a non-finite coordinate are left as they are.  T: column-major 4x4.
struct SegPose {
  float m[16];
};
__global__ void __launch_bounds__(256) k_seg_transform(int n, const float* __restrict__ scan, int stride, SegPose T, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* q = scan + (size_t)i * stride;
  const float x = q[0], y = q[1], z = q[2];
  float4 o = make_float4(x, y, z, 1.0f);
  if (isfinite(x) && isfinite(y) && isfinite(z)) {
    o.x = __fadd_rn(__fadd_rn(__fmul_rn(T.m[0], x), __fmul_rn(T.m[4], y)), __fadd_rn(__fmul_rn(T.m[8], z), T.m[12]));
    o.y = __fadd_rn(__fadd_rn(__fmul_rn(T.m[1], x), __fmul_rn(T.m[5], y)), __fadd_rn(__fmul_rn(T.m[9], z), T.m[13]));
    o.z = __fadd_rn(__fadd_rn(__fmul_rn(T.m[2], x), __fmul_rn(T.m[6], y)), __fadd_rn(__fmul_rn(T.m[10], z), T.m[14]));
  }
  out[i] = o;
}

// (also clears the per-pixel work arrays of the later kernels: one launch instead of five memsets)
__global__ void __launch_bounds__(256) k_seg_project(SegDev p, const float* __restrict__ scan, int stride, float* __restrict__ range,
                                                     signed char* __restrict__ ground, int* __restrict__ label0, int* __restrict__ size,
                                                     int* __restrict__ accepted, unsigned* __restrict__ cstat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.HW) return;
  size[i] = 0;
  accepted[i] = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) cstat[(size_t)k * p.HW + i] = (k == 1 || k == 4) ? 0xffffffffu : 0u;  // "none" of the min / max statistics
  const int row = i / p.W, col = i - row * p.W;
  float x, y, z, r;
  seg_full_point(scan, stride, p, i, x, y, z, r);
  // The column loop runs row = H-1 down to H-ground_rows; iteration `row` may write rows `row` and `row - 1`.
  // What row r ends with: iteration r+1 first (ground -> 1), then iteration r (no info -> -1, ground -> 1).
  const int first = p.H - p.ground_rows;  // smallest row that has an iteration of its own
  int g = 0;
  if (row + 1 >= first && row + 1 <= p.H - 1 && seg_pair_state(scan, stride, p, row + 1, col) == 1) g = 1;
  if (row >= first) {
    const int s = seg_pair_state(scan, stride, p, row, col);
    if (s != 0) g = s;
  }
  range[i] = r;
  ground[i] = (signed char)g;
  label0[i] = (g == 1 || r == 0.0f) ? -1 : 0;  // :497-508
}

// bit n of a pixel's nibble: the flood fill standing on this pixel would push neighbour n (:577-636), given
// that the neighbour has not been labelled by then.  n: 0 up, 1 right, 2 left, 3 down (neighbor_iterator_, :132-147).
__device__ unsigned seg_edge_bits(const SegDev& p, const float* __restrict__ range, const int* __restrict__ label0, int i) {
  const int y = i / p.W, x = i - y * p.W;
  if (label0[i] != 0 || !seg_in_window(p, y, x)) return 0u;
  const float rf = range[i];
  unsigned bits = 0u;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int dy = n == 0 ? -1 : (n == 3 ? 1 : 0), dx = n == 1 ? 1 : (n == 2 ? -1 : 0);
    const int ty = y + dy, tx = x + dx;
    if (ty < 0 || ty >= p.H || tx < 0 || tx >= p.W) continue;  // tx = -1 fails the reference's unsigned window test; tx = W needs wc1 >= W
    if (!seg_in_window(p, ty, tx)) continue;
    const int t = ty * p.W + tx;
    if (label0[t] != 0) continue;
    const float rt = range[t];
    const float d1 = fmaxf(rf, rt), d2 = fminf(rf, rt);
    const float sa = dy == 0 ? p.sin_x : p.sin_y, ca = dy == 0 ? p.cos_x : p.cos_y;
    const float angle = (float)atan2((double)__fmul_rn(d2, sa), (double)__fsub_rn(d1, __fmul_rn(d2, ca)));
    if (angle > p.theta) bits |= 1u << n;
  }
  return bits;
}

// One pixel per thread: the edge nibbles (two pixels per byte) and the start of the union-find.  Pixels joined by
// left/right edges form runs of consecutive raster indices (edge bits are clear at the row ends); inside a warp's 32
// pixels every pixel is hung directly under the first pixel of its run (or of the warp's part of it), so only the
// links between warps and between rows are left for k_ccl_union.
__global__ void __launch_bounds__(256) k_seg_edges(SegDev p, const float* __restrict__ range, const int* __restrict__ label0,
                                                   unsigned char* __restrict__ nib, int* __restrict__ parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  unsigned b = 0u;
  bool ok = false;
  if (i < p.HW) {
    const int y = i / p.W, x = i - y * p.W;
    ok = label0[i] == 0 && seg_in_window(p, y, x);
    b = seg_edge_bits(p, range, label0, i);
  }
  const unsigned hi = __shfl_down_sync(kFullMask, b, 1);
  if (i < p.HW && (lane & 1) == 0) nib[i >> 1] = (unsigned char)(b | (hi << 4));
  const unsigned starts = __ballot_sync(kFullMask, ok && (b & 4u) == 0u);  // pixels without a left edge open a run
  if (i < p.HW) {
    const unsigned upto = starts & (0xffffffffu >> (31 - lane));
    parent[i] = !ok ? -1 : (i - lane + (upto ? 31 - __clz(upto) : 0));
  }
}

__device__ __forceinline__ int ccl_find(const int* parent, int x) {
  int q;
  while ((q = __ldcg(parent + x)) != x) x = q;
  return x;
}
__device__ void ccl_union(int* parent, int a, int b) {
  while (true) {
    a = ccl_find(parent, a);
    b = ccl_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b, b = t;
    }
    const int old = atomicMin(parent + a, b);  // hang the larger root under the smaller
    if (old == a) return;
    a = old;  // somebody moved a meanwhile: merge what it points to now
  }
}
__device__ __forceinline__ unsigned seg_nibble(const unsigned char* __restrict__ nib, int i) { return ((unsigned)nib[i >> 1] >> ((i & 1) * 4)) & 15u; }
__global__ void __launch_bounds__(256) k_ccl_union(SegDev p, const unsigned char* __restrict__ nib, int* __restrict__ parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.HW) return;
  const unsigned b = seg_nibble(nib, i);
  // a run that continues across the warp boundary
  if ((threadIdx.x & 31) == 0 && (b & 4u)) ccl_union(parent, i, i - 1);
  // a link to the row below, unless the left neighbour makes the same link between the same two runs
  if (b & 8u) {
    const bool repeated = (b & 4u) && (seg_nibble(nib, i - 1) & 8u) && (seg_nibble(nib, i + p.W) & 4u);
    if (!repeated) ccl_union(parent, i, i + p.W);
  }
}
// order-preserving map float -> unsigned (and back), for atomicMin / atomicMax on floats of either sign
__device__ __forceinline__ unsigned seg_f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u ^ 0x80000000u);
}
__device__ __forceinline__ float seg_ord2f(unsigned o) { return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o); }

// Per-segment statistics that do not depend on the flood fill's order, cstat[k * HW + root]:
//   0 largest range of any pixel (float bits; ranges are positive)      3 largest z of the pushed pixels (ordered, 0 = none)
//   1 smallest row of the pushed pixels (0xffffffff = none)             4 smallest non-zero z of the pushed pixels (ordered,
//   2 largest row + 1 of the pushed pixels (0 = none)                     0xffffffff = none)
//   5 largest z of the pushed pixels other than the FIRST pushed one (ordered, 0 = none)
// "pushed" = every pixel of the segment but its seed (the seed is popped, never pushed: :556-569); the first pushed pixel
// is the seed's first neighbour in direction order.
constexpr int kStatRange = 0, kStatRowMin = 1, kStatRowMax = 2, kStatZMax = 3, kStatZMin = 4, kStatZMaxRest = 5, kStatCount = 6;
__device__ __forceinline__ int seg_step(int d, int W) { return d == 0 ? -W : (d == 1 ? 1 : (d == 2 ? -1 : W)); }

// root[] = smallest raster index of the pixel's component (-1 outside), size[root] = pixels of the component, and the
// statistics above, folded per warp over the lanes that share a root before they go to memory
__global__ void __launch_bounds__(256) k_ccl_flatten(SegDev p, const int* __restrict__ parent, const float* __restrict__ scan, int stride,
                                                     const float* __restrict__ range, const unsigned char* __restrict__ nib,
                                                     int* __restrict__ root, int* __restrict__ size, unsigned* __restrict__ cstat,
                                                     const float* __restrict__ residuals, int res_stride, double* __restrict__ res_sum,
                                                     int* __restrict__ res_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int r = -1;
  if (i < p.HW && parent[i] >= 0) r = ccl_find(parent, i);
  if (i < p.HW) root[i] = r;
  const bool member = r >= 0, pushed = member && r != i;
  const float z = pushed ? scan[(size_t)i * stride + 2] : 0.0f;
  const unsigned row = (unsigned)(i / p.W);
  bool first_pushed = false;
  if (pushed) {
    const unsigned nb = seg_nibble(nib, r);  // not empty: the seed has company
    first_pushed = i == r + seg_step(__ffs(nb) - 1, p.W);
  }
  const unsigned m = __match_any_sync(kFullMask, r);
  const unsigned s_range = __reduce_max_sync(m, member ? __float_as_uint(range[i]) : 0u);
  const unsigned s_rmin = __reduce_min_sync(m, pushed ? row : 0xffffffffu);
  const unsigned s_rmax = __reduce_max_sync(m, pushed ? row + 1u : 0u);
  const unsigned s_zmax = __reduce_max_sync(m, pushed ? seg_f2ord(z) : 0u);
  const unsigned s_zmin = __reduce_min_sync(m, (pushed && z != 0.0f) ? seg_f2ord(z) : 0xffffffffu);
  const unsigned s_zrest = __reduce_max_sync(m, (pushed && !first_pushed) ? seg_f2ord(z) : 0u);
  if (member && (int)(__ffs(m) - 1) == (int)(threadIdx.x & 31)) {
    atomicAdd(size + r, __popc(m));
    atomicMax(cstat + (size_t)kStatRange * p.HW + r, s_range);
    atomicMin(cstat + (size_t)kStatRowMin * p.HW + r, s_rmin);
    atomicMax(cstat + (size_t)kStatRowMax * p.HW + r, s_rmax);
    atomicMax(cstat + (size_t)kStatZMax * p.HW + r, s_zmax);
    atomicMin(cstat + (size_t)kStatZMin * p.HW + r, s_zmin);
    atomicMax(cstat + (size_t)kStatZMaxRest * p.HW + r, s_zrest);
  }
  if (p.unordered_sums && p.have_residuals && pushed) {  // opt-in: total_residuum without the push order (:632-634)
    const float rv = residuals[(size_t)i * res_stride];
    if (rv > 0.0f) {
      atomicAdd(res_sum + r, (double)rv);
      atomicAdd(res_count + r, 1);
    }
  }
}
// per pixel (size << 32 | 1) at the seeds, 0 elsewhere; its exclusive sum gives every seed the start of its push list
// (high word) and its ordinal among the seeds in raster order (low word)
__global__ void __launch_bounds__(256) k_seg_seed_keys(SegDev p, const int* __restrict__ root, const int* __restrict__ size,
                                                       unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.HW) return;
  keys[i] = root[i] == i ? (((unsigned long long)size[i] << 32) | 1ull) : 0ull;
}
__global__ void __launch_bounds__(256) k_seg_seeds(SegDev p, const int* __restrict__ root, const unsigned long long* __restrict__ pre,
                                                   int* __restrict__ seeds, int* __restrict__ n_seeds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.HW) return;
  const bool is_seed = root[i] == i;
  const int ordinal = (int)(pre[i] & 0xffffffffull);
  if (is_seed) seeds[ordinal] = i;
  if (i == p.HW - 1) *n_seeds = ordinal + (is_seed ? 1 : 0);
}

// The queue flood fill of labelComponents, one warp per segment, and the segment tests (:548-724).
__global__ void __launch_bounds__(kFillThreads) k_seg_fill(SegDev p, const unsigned char* __restrict__ nib, const int* __restrict__ seeds,
                                                           const int* __restrict__ n_seeds, const unsigned long long* __restrict__ pre,
                                                           const float* __restrict__ scan, int stride, const float* __restrict__ range,
                                                           const float* __restrict__ residuals, int res_stride, unsigned* __restrict__ order,
                                                           int* __restrict__ next_seed, int* __restrict__ accepted, double* __restrict__ seg_avg,
                                                           int ring_size, const int* __restrict__ size, const unsigned* __restrict__ cstat,
                                                           int shortcut, const double* __restrict__ res_sum, const int* __restrict__ res_count) {
  extern __shared__ __align__(16) unsigned char seg_smem[];
  const int nib_bytes = ((p.HW + 1) / 2 + 15) & ~15;
  const int vis_words = (p.HW + 31) / 32;
  unsigned char* snib = seg_smem;
  unsigned* vis = reinterpret_cast<unsigned*>(seg_smem + nib_bytes);
  int* rings = reinterpret_cast<int*>(vis + ((vis_words + 3) & ~3));
  for (int i = threadIdx.x; i < nib_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(snib)[i] = reinterpret_cast<const uint4*>(nib)[i];
  for (int i = threadIdx.x; i < vis_words; i += blockDim.x) vis[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pl = lane >> 2, n = lane & 3;  // flood fill: slot = (queue entry of the step, direction)
  const int step_n = n == 0 ? -p.W : (n == 1 ? 1 : (n == 2 ? -1 : p.W));
  const unsigned lanes_below = (1u << lane) - 1u;
  int* ring = rings + warp * kRing;
  const int total = *n_seeds;
  const int W = p.W;
  while (true) {
    int si = 0;
    if (lane == 0) si = atomicAdd(next_seed, 1);
    si = __shfl_sync(kFullMask, si, 0);
    if (si >= total) break;
    const int seed = seeds[si];
    const unsigned beg = (unsigned)(pre[seed] >> 32);
    // ---- Most segments are decided without replaying the queue.  Of the segment tests (:640-685) only max_z depends on the
    // push order: it is the largest z among the pushed pixels that did not set a new minimum when they were pushed
    // (:612-616).  A pixel sets a new (strict) minimum only if every non-zero z pushed before it is larger, so
    //   * the highest pixel counts unless it is the first pushed pixel p1 (the seed's first neighbour in direction order);
    //   * if p1 is the highest, the highest of the others counts if it ties with p1, or if it is not the second pushed
    //     pixel p2 (p2 came before it with a smaller z);
    // anything else (a zero z in p1 or p2, p2 the unique runner-up) is left to the replay.  Size, scan lines, max_dist (= the
    // largest range of the segment once anything was pushed: the seed enters through its children), min_z and the
    // elevation never depend on the order.  A segment that fails a test is rejected here; one that passes still needs the
    // replay when residuals are summed in push order.
    if (shortcut) {
      const int sz = size[seed];
      const unsigned rmin = cstat[(size_t)kStatRowMin * p.HW + seed], rmax = cstat[(size_t)kStatRowMax * p.HW + seed];
      const unsigned ozmax = cstat[(size_t)kStatZMax * p.HW + seed], ozmin = cstat[(size_t)kStatZMin * p.HW + seed];
      const unsigned ozrest = cstat[(size_t)kStatZMaxRest * p.HW + seed];
      const int lines = rmax > rmin ? (int)(rmax - rmin) : 0;
      const float max_dist = sz >= 2 ? __uint_as_float(cstat[(size_t)kStatRange * p.HW + seed]) : -1e6f;
      const float min_z = ozmin == 0xffffffffu ? 1e6f : seg_ord2f(ozmin);
      const float top_z = ozmax == 0u ? -1e6f : seg_ord2f(ozmax);
      bool pass = (sz >= 50 && lines >= p.min_line_num) || (sz >= p.valid_point_num && lines >= p.valid_line_num);
      pass = pass && max_dist <= p.max_distance && __fsub_rn(min_z, p.height) <= p.max_elevation;
      bool decided = !pass;
      const unsigned nb = ((unsigned)snib[seed >> 1] >> ((seed & 1) * 4)) & 15u;
      if (pass && nb != 0u) {
        const int d1 = __ffs(nb) - 1;
        const int p1 = seed + seg_step(d1, W);
        const float z1 = scan[(size_t)p1 * stride + 2];
        float max_z = 0.0f;
        bool known = false;
        if (z1 != 0.0f && z1 != top_z) {
          max_z = top_z, known = true;
        } else if (z1 != 0.0f) {  // p1 is the highest pixel and sets the first minimum: it never counts
          if (ozrest == 0u) {
            max_z = -1e6f, known = true;  // nothing else was pushed
          } else {
            const float t = seg_ord2f(ozrest);
            if (t == z1) {
              max_z = t, known = true;  // as high as p1 and pushed later: not a strict minimum
            } else {
              const unsigned nb_rest = nb & (nb - 1u);
              int p2 = -1;
              if (nb_rest != 0u) {
                p2 = seed + seg_step(__ffs(nb_rest) - 1, W);  // the seed's second neighbour
              } else {
                const unsigned nb1 = (((unsigned)snib[p1 >> 1] >> ((p1 & 1) * 4)) & 15u) & ~(1u << (3 - d1));  // p1's neighbours but the seed
                if (nb1 != 0u) p2 = p1 + seg_step(__ffs(nb1) - 1, W);
              }
              if (p2 >= 0) {
                const float z2 = scan[(size_t)p2 * stride + 2];
                if (z2 != 0.0f && z2 != t) max_z = t, known = true;  // p2 came first with a smaller z: the runner-up counts
              }
            }
          }
        }
        if (known) {
          const float delta_z = __fsub_rn(max_z, min_z);
          pass = p.min_delta_z <= delta_z && delta_z <= p.max_delta_z;
          decided = !pass || !p.have_residuals || p.unordered_sums;
        }
      }
      if (decided) {
        if (lane == 0) {
          accepted[si] = pass ? 1 : 0;
          double avg = 0.0;
          if (pass && p.have_residuals && p.unordered_sums && res_count[seed] > 0)
            avg = (double)__fdiv_rn((float)res_sum[seed], (float)res_count[seed]);  // float total / count, as :697
          seg_avg[si] = avg;
        }
        continue;
      }
    }
    // ---- the queue: entries [head, tail) wait to be popped; entry 0 is the seed
    int head = 0, tail = 1;
    if (lane == 0) {
      ring[0] = seed;
      order[beg] = (unsigned)seed;
      atomicOr(vis + (seed >> 5), 1u << (seed & 31));
    }
    __syncwarp();
    while (head < tail) {
      const int cnt = min(8, tail - head);
      bool cand = false;
      int v = 0;
      if (pl < cnt) {
        const int q = head + pl;
        const int u = (tail - q <= ring_size) ? ring[q & (ring_size - 1)] : (int)(__ldcg(order + beg + q) & 0x3fffffffu);
        if (((unsigned)snib[u >> 1] >> ((u & 1) * 4 + n)) & 1u) {
          v = u + step_n;
          cand = ((vis[v >> 5] >> (v & 31)) & 1u) == 0u;
        }
      }
      // A pixel wanted by several (entry, direction) slots of this step goes to the first of them.  (The match costs
      // one round per distinct value: every slot without a candidate shares one dummy.)
      const unsigned same = __match_any_sync(kFullMask, cand ? v : -1);
      const bool win = cand && (same & lanes_below) == 0u;
      const unsigned wins = __ballot_sync(kFullMask, win);
      if (win) {
        const int pos = tail + __popc(wins & lanes_below);
        ring[pos & (ring_size - 1)] = v;
        order[beg + pos] = (unsigned)v | ((unsigned)n << 30);  // the direction recovers the pixel it was pushed from
        atomicOr(vis + (v >> 5), 1u << (v & 31));
      }
      tail += __popc(wins);
      head += cnt;
      __syncwarp();
    }
    // ---- statistics over the pushed pixels in push order (the seed itself only counts for the size, :556-569).
    // The gathers of a 32-pixel group are issued one group ahead of their use.
    float min_z = 1e6f, max_z = -1e6f, max_dist = -1e6f, total_res = 0.0f;
    int res_count = 0, row_lo = 0x7fffffff, row_hi = -1;
    auto fetch = [&](int base) -> unsigned {
      const int i = base + lane;
      return i < tail ? __ldcg(order + beg + i) : 0xffffffffu;
    };
    struct Pushed {
      float z, r, d;
      int row;
      bool active;
    };
    auto gather = [&](unsigned e) -> Pushed {
      Pushed g{0.0f, 0.0f, -1e6f, 0, e != 0xffffffffu};
      if (g.active) {
        const int v = (int)(e & 0x3fffffffu), dir = (int)(e >> 30);
        const int u = v - (dir == 0 ? -W : (dir == 1 ? 1 : (dir == 2 ? -1 : W)));
        g.d = fmaxf(range[u], range[v]);
        g.z = scan[(size_t)v * stride + 2];
        if (p.have_residuals) g.r = residuals[(size_t)v * res_stride];
        g.row = v / W;
      }
      return g;
    };
    Pushed cur = gather(fetch(1));
    unsigned e_ahead = fetch(33);
    for (int base = 1; base < tail; base += 32) {
      const Pushed nxt = gather(e_ahead);
      e_ahead = fetch(base + 64);
      const bool active = cur.active;
      const float z = cur.z;
      max_dist = fmaxf(max_dist, cur.d);
      if (active) row_lo = min(row_lo, cur.row), row_hi = max(row_hi, cur.row);
      // `if (z < min_z && z != 0) min_z = z; else if (z > max_z) max_z = z;` (:612-616): min_z before pixel i is the
      // minimum of the non-zero z pushed before it
      const float val = (active && z != 0.0f && !isnan(z)) ? z : INFINITY;
      float incl = val;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl = fminf(incl, t);
      }
      float before = __shfl_up_sync(kFullMask, incl, 1);
      if (lane == 0) before = INFINITY;
      before = fminf(min_z, before);
      const bool new_min = active && z != 0.0f && z < before;
      if (active && !new_min && z > max_z) max_z = z;
      min_z = fminf(min_z, __shfl_sync(kFullMask, incl, 31));
      // total_residuum += residual, in push order (:632-634): a lane-ordered chain; the zeros of skipped pixels and of
      // the lanes behind the end change nothing
      const bool counted = active && cur.r > 0.0f;
      res_count += counted ? 1 : 0;
      const float rv = counted ? cur.r : 0.0f;
      if (__any_sync(kFullMask, counted)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) total_res = __fadd_rn(total_res, __shfl_sync(kFullMask, rv, j));
      }
      cur = nxt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      max_z = fmaxf(max_z, __shfl_xor_sync(kFullMask, max_z, o));
      max_dist = fmaxf(max_dist, __shfl_xor_sync(kFullMask, max_dist, o));
      res_count += __shfl_xor_sync(kFullMask, res_count, o);
      row_lo = min(row_lo, __shfl_xor_sync(kFullMask, row_lo, o));
      row_hi = max(row_hi, __shfl_xor_sync(kFullMask, row_hi, o));
    }
    // line_count (:640-644): the rows of a 4-connected segment are consecutive, and so are those of its pushed pixels
    // (the seed lies in the top row; without it that row either stays or goes as a whole)
    const int line_count = row_hi >= row_lo ? row_hi - row_lo + 1 : 0;
    // ---- the segment tests (:640-685)
    bool feasible = false;
    if (tail >= 50 && line_count >= p.min_line_num)
      feasible = true;
    else if (tail >= p.valid_point_num && line_count >= p.valid_line_num)
      feasible = true;
    if (feasible) feasible = max_dist <= p.max_distance;
    if (feasible) {
      const float delta_z = __fsub_rn(max_z, min_z);
      feasible = p.min_delta_z <= delta_z && delta_z <= p.max_delta_z;
    }
    if (feasible) feasible = __fsub_rn(min_z, p.height) <= p.max_elevation;
    if (lane == 0) {
      accepted[si] = feasible ? 1 : 0;
      seg_avg[si] = (p.have_residuals && res_count > 0) ? (double)__fdiv_rn(total_res, (float)res_count) : 0.0;
    }
    __syncwarp();
  }
}

// accepted segments get 1 + their rank among the accepted ones in seed order; the others 999999 (:687-722)
__global__ void __launch_bounds__(256) k_seg_labels(SegDev p, const int* __restrict__ root, const unsigned long long* __restrict__ pre,
                                                    const int* __restrict__ accepted, const int* __restrict__ rank,
                                                    const double* __restrict__ seg_avg, const int* __restrict__ n_seeds, int* __restrict__ label,
                                                    double* __restrict__ avg_by_label, int* __restrict__ label_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    const int m = *n_seeds;
    *label_count = m > 0 ? 1 + rank[m - 1] + accepted[m - 1] : 1;
  }
  if (i >= p.HW) return;
  const int r = root[i];
  if (r < 0) return;  // keeps the initial label: -1 ground / empty, 0 outside the window
  const int ordinal = (int)(pre[r] & 0xffffffffull);
  const bool ok = accepted[ordinal] != 0;
  const int lab = 1 + rank[ordinal];
  label[i] = ok ? lab : 999999;
  if (ok && r == i) avg_by_label[lab] = seg_avg[ordinal];
}

// the stage's temporaries: stream-ordered pool buffers, all returned on every exit path
struct SegPool {
  ddlo_runtime* rt;
  std::vector<void*> held;
  explicit SegPool(ddlo_runtime* r) : rt(r) {}
  ~SegPool() {
    for (void* q : held) cudaFreeAsync(q, rt->stream);
  }
  template <class T>
  int alloc(T** ptr, size_t count) {
    DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(ptr), std::max<size_t>(count, 1) * sizeof(T), rt->stream));
    held.push_back(*ptr);
    return DDLO_OK;
  }
};

}  // namespace

// All pointers are device pointers; d_label doubles as the initial label image; the residual of pixel i is
// d_residuals[i * res_stride] (1 for a plane, 4 for the w channel of a residual image).  Enqueues only.
int segment_scan_device(ddlo_runtime* rt, const ddlo_segmentation_params& prm, const float* T16, const float* d_scan, int stride_floats,
                        const float* d_residuals, int res_stride, int* d_label, float* d_range, signed char* d_ground, double* d_avg_by_label,
                        int* d_label_count) {
  SegDev p{};
  p.H = prm.rows, p.W = prm.cols, p.HW = prm.rows * prm.cols;
  p.ground_rows = prm.ground_rows;
  p.wr0 = std::max(prm.window_row_min, 0), p.wr1 = prm.window_row_max, p.wc0 = std::max(prm.window_col_min, 0), p.wc1 = prm.window_col_max;
  p.valid_point_num = prm.valid_point_num, p.min_line_num = prm.min_line_num, p.valid_line_num = prm.valid_line_num;
  p.minimum_range = prm.minimum_range, p.ground_thr = prm.ground_angle_threshold, p.mount = prm.sensor_mount_angle, p.theta = prm.theta;
  // loadParams (:78-79, :109-112): float resolutions, double sin/cos, float members
  const float ang_res_x = 360.0 / float(prm.cols);
  const float ang_res_y = 2 * prm.ang_bottom / float(prm.rows - 1);
  p.sin_x = std::sin(ang_res_x / 180.0 * M_PI), p.cos_x = std::cos(ang_res_x / 180.0 * M_PI);
  p.sin_y = std::sin(ang_res_y / 180.0 * M_PI), p.cos_y = std::cos(ang_res_y / 180.0 * M_PI);
  p.x0 = -T16[12], p.y0 = -T16[13], p.z0 = -T16[14], p.height = T16[14];
  p.min_delta_z = prm.min_delta_z, p.max_delta_z = prm.max_delta_z, p.max_distance = prm.max_distance, p.max_elevation = prm.max_elevation;
  p.have_residuals = d_residuals ? 1 : 0;
  p.unordered_sums = prm.unordered_residual_sums ? 1 : 0;

  const int HW = p.HW;
  const size_t nib_bytes = (((size_t)HW + 1) / 2 + 15) & ~(size_t)15;
  const size_t vis_words = (((size_t)HW + 31) / 32 + 3) & ~(size_t)3;
  const size_t smem = nib_bytes + vis_words * 4 + (size_t)kFillWarps * kRing * 4;
  if (smem > 227 * 1024) return fail(DDLO_E_UNSUPPORTED, "range image too large for the shared-memory flood fill (rows * cols <= ~300000)");
  static std::atomic<unsigned long long> configured{0};
  if (!((configured.load() >> rt->device) & 1ull)) {
    DDLO_CUDA(cudaFuncSetAttribute(k_seg_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.fetch_or(1ull << rt->device);
  }

  cudaStream_t st = rt->stream;
  unsigned char* nib = nullptr;
  int *parent = nullptr, *root = nullptr, *size = nullptr, *seeds = nullptr, *accepted = nullptr, *rank = nullptr, *small = nullptr;
  unsigned long long* keys = nullptr;
  unsigned *order = nullptr, *cstat = nullptr;
  double *seg_avg = nullptr, *res_sum = nullptr;
  int* res_count = nullptr;
  void* tmp = nullptr;
  const size_t tmp_bytes = scan_temp_bytes((size_t)HW);
  SegPool pool(rt);
  DDLO_TRY(pool.alloc(&nib, nib_bytes));
  DDLO_TRY(pool.alloc(&parent, HW));
  DDLO_TRY(pool.alloc(&root, HW));
  DDLO_TRY(pool.alloc(&size, HW));
  DDLO_TRY(pool.alloc(&seeds, HW));
  DDLO_TRY(pool.alloc(&accepted, HW));
  DDLO_TRY(pool.alloc(&rank, HW));
  DDLO_TRY(pool.alloc(&small, 4));  // [0] seeds found, [1] next seed to fill
  DDLO_TRY(pool.alloc(&keys, HW));
  DDLO_TRY(pool.alloc(&order, HW));
  DDLO_TRY(pool.alloc(&cstat, (size_t)kStatCount * HW));
  if (p.unordered_sums && p.have_residuals) {
    DDLO_TRY(pool.alloc(&res_sum, HW));
    DDLO_TRY(pool.alloc(&res_count, HW));
    DDLO_CUDA(cudaMemsetAsync(res_sum, 0, (size_t)HW * 8, st));
    DDLO_CUDA(cudaMemsetAsync(res_count, 0, (size_t)HW * 4, st));
  }
  DDLO_TRY(pool.alloc(&seg_avg, HW));
  DDLO_TRY(pool.alloc(reinterpret_cast<unsigned char**>(&tmp), tmp_bytes));
  DDLO_CUDA(cudaMemsetAsync(small, 0, 16, st));
  DDLO_CUDA(cudaMemsetAsync(nib + nib_bytes - 16, 0, 16, st));  // the padding behind the last pixel pair

  const int pb = (HW + 255) / 256;
  float4* moved = nullptr;
  if (prm.scan_in_sensor_frame) {
    DDLO_TRY(pool.alloc(&moved, HW));
    SegPose pose;
    for (int i = 0; i < 16; ++i) pose.m[i] = T16[i];
    k_seg_transform<<<pb, 256, 0, st>>>(HW, d_scan, stride_floats, pose, moved);
    rt->launches += 1;
    d_scan = reinterpret_cast<const float*>(moved);
    stride_floats = 4;
  }
  static_assert(kStatCount == 6 && kStatRowMin == 1 && kStatZMin == 4, "k_seg_project initialises the statistics by index");
  k_seg_project<<<pb, 256, 0, st>>>(p, d_scan, stride_floats, d_range, d_ground, d_label, size, accepted, cstat);
  k_seg_edges<<<pb, 256, 0, st>>>(p, d_range, d_label, nib, parent);
  k_ccl_union<<<pb, 256, 0, st>>>(p, nib, parent);
  k_ccl_flatten<<<pb, 256, 0, st>>>(p, parent, d_scan, stride_floats, d_range, nib, root, size, cstat, d_residuals, res_stride, res_sum, res_count);
  k_seg_seed_keys<<<pb, 256, 0, st>>>(p, root, size, keys);
  DDLO_TRY(scan_u64(st, keys, keys, (size_t)HW, false, tmp, nullptr));
  k_seg_seeds<<<pb, 256, 0, st>>>(p, root, keys, seeds, small);
  const int per_sm = smem <= 100 * 1024 ? 2 : 1;
  int ring_size = kRing;  // testing aid: a small ring forces the queue reads through global memory
  if (const char* e = std::getenv("DDLO_SEG_RING")) {
    const int v = std::atoi(e);
    if (v >= 64 && v <= kRing && (v & (v - 1)) == 0) ring_size = v;
  }
  const int shortcut = std::getenv("DDLO_SEG_NO_SHORTCUT") ? 0 : 1;  // testing aid: replay the queue of every segment
  k_seg_fill<<<rt->num_sms * per_sm, kFillThreads, smem, st>>>(p, nib, seeds, small, keys, d_scan, stride_floats, d_range, d_residuals, res_stride, order,
                                                              small + 1, accepted, seg_avg, ring_size, size, cstat, shortcut, res_sum, res_count);
  DDLO_TRY(scan_int(st, accepted, rank, (size_t)HW, false, tmp, nullptr));
  k_seg_labels<<<pb, 256, 0, st>>>(p, root, keys, accepted, rank, seg_avg, small, d_label, d_avg_by_label, d_label_count);
  rt->launches += 8 + 6;  // + the two scans (three kernels each, prims.cu)
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;  // ~SegPool returns the temporaries
}

}  // namespace ddlo
