// Shared declarations of the B200 nano_gicp library (host + device).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/ddlo_gicp.h"

// -DDDLO_BOUNDS_CHECK builds (compute-sanitizer is not available on every pool): every gather / scatter index of the
// search, linearize, covariance and sort kernels is checked on the device; the first violation prints where and traps.
#ifdef DDLO_BOUNDS_CHECK
#define DDLO_CHECK_INDEX(i, n, what)                                                                                          \
  do {                                                                                                                        \
    if ((unsigned long long)(long long)(i) >= (unsigned long long)(long long)(n)) {                                           \
      printf("ddlo bounds check: %s: index %lld not in [0, %lld) (block %d, thread %d)\n", what, (long long)(i), (long long)(n), \
             (int)blockIdx.x, (int)threadIdx.x);                                                                              \
      __trap();                                                                                                               \
    }                                                                                                                         \
  } while (0)
#else
#define DDLO_CHECK_INDEX(i, n, what) \
  do {                               \
  } while (0)
#endif

namespace ddlo {

// ---------------------------------------------------------------------------------------------
// error plumbing: thread-local message, int status across the C boundary
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define DDLO_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return ::ddlo::fail(DDLO_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
  } while (0)

#define DDLO_TRY(expr)            \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != DDLO_OK) return rc__; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// kNN index: Morton-ordered points + an explicit octree over Morton prefixes.
//
//   spts[i]   i-th point in Morton order (30-bit code, 10 bits per axis on one isotropic lattice over
//             the cloud's bounding box); .w carries the ORIGINAL index (int bits).
//   cell      the set of points sharing the first 3t bits of the code (level t, t = 0..10); cells of
//             one level are disjoint cubes, so the tight boxes of sibling cells are disjoint too.
//   leaf      the largest cell holding at most kLeafMax points (level-10 cells are leaves whatever
//             their size); a contiguous run spts[start, start+count).
//   node      every cell above a leaf; 256 bytes = 16 float4:
//               [0..11]  boxes of the 8 child cells, SoA: lo.x x8, lo.y x8, lo.z x8, hi.x x8, hi.y x8, hi.z x8
//               [12..15] 8 child references (int2): {node id, -1} internal, {start, count>0} leaf, {0,0} empty
//             empty slots keep lo=+inf, hi=-inf, whose distance bound is +inf and never qualifies.
//   node 0 is the root (level 0, the whole cloud); ids are breadth-first.
//   meta / node_of_point let a search that already knows a nearby point of the cloud start deep in
//   the tree instead of at the root (knn.cuh, ball_in_cell).
// ---------------------------------------------------------------------------------------------
#ifndef DDLO_LEAF_MAX
#define DDLO_LEAF_MAX 16
#endif
constexpr int kLeafMax = DDLO_LEAF_MAX;
constexpr int kMortonLevels = 10;
constexpr int kNodeF4 = 16;     // float4 per node
constexpr int kStackDepth = 80; // pending internal children: <= 7 per level, <= 11 levels

struct IndexView {
  const float4* spts;
  const float4* nodes;
  const int4* meta;          // per node: {parent id (-1 root), level, cell origin on the 10-bit lattice (x | y<<10 | z<<20), 0}
  const int* node_of_point;  // per ORIGINAL point index: the node whose leaf child holds the point
  const float* lattice;      // device: {lo.x, lo.y, lo.z, scale} of the Morton lattice, u = (p - lo) * scale
  int n;
};

// 6 doubles per point: xx, xy, xz, yy, yz, zz (the 3x3 block of the reference's Matrix4d)
constexpr int kCovStride = 6;

// number of fp64 sums one linearize produces: 21 (upper H) + 6 (b) + 1 (error)
constexpr int kNumSums = 28;

struct RuntimeImpl;

}  // namespace ddlo

// opaque handle bodies ---------------------------------------------------------------------------
struct ddlo_runtime {
  // Reference counted: the creator holds one reference (dropped by ddlo_runtime_destroy), every cloud, covariance
  // vector and engine made from the runtime holds another, so a handle that outlives the destroy call - e.g. a
  // shared cloud still set as another runtime's target - never dangles; the stream goes with the last handle.
  std::atomic<int> refs{1};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;  // short independent work inside one call (node array initialisation during the sort)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t side2 = nullptr;  // the first correspondence search of an align that computes covariances first
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t slots[16] = {};
  int num_sms = 0;
  long long launches = 0;
  void* flush_buf = nullptr;
  size_t flush_bytes = 0;
  // pinned scratch for small read-backs
  void* h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  // device scratch reused by the index build (radix sort temp etc.)
  void* d_scratch = nullptr;
  size_t d_scratch_bytes = 0;
  int max_coop_blocks_align = 0;  // co-resident blocks of the align kernel (one 1024-thread block per SM)
  int align_blocks_limit = 0;     // blocks one align may use (ddlo_runtime_set_align_blocks); default = all
};

struct ddlo_cloud {
  std::atomic<int> refs{1};
  ddlo_runtime* rt = nullptr;
  int n = 0;
  float4* pts = nullptr;  // original order, w = 1
  // index (optional)
  bool has_index = false;
  float4* spts = nullptr;   // Morton order, w = original index
  float4* nodes = nullptr;  // octree nodes, kNodeF4 float4 each
  int4* meta = nullptr;
  int* node_of_point = nullptr;
  float* lattice = nullptr;  // 4 floats + the node count (int) behind them
  ddlo::IndexView view{};
  // ddlo_cloud_share: the handle's work is complete and synchronised; engines of other runtimes (streams) of the
  // same device may read it, nobody may add an index to it any more
  bool shared = false;
};

struct ddlo_covs {
  std::atomic<int> refs{1};
  ddlo_runtime* rt = nullptr;
  int n = 0;
  double* c = nullptr;  // n * 6
  // the same covariances permuted into the Morton order of one cloud's index (target role: the align kernel gathers
  // the matched target covariance by Morton position); cached here, keyed by the retained cloud handle
  double* sorted = nullptr;
  ddlo_cloud* sorted_for = nullptr;
  bool shared = false;  // see ddlo_cloud::shared
};
