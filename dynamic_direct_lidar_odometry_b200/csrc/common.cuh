// Shared declarations of the B200 nano_gicp library (host + device).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/ddlo_gicp.h"

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
// kNN index: Morton-ordered points + an implicit 8-wide tree of axis-aligned boxes.
//
//   spts[i]           i-th point in Morton order, .w carries the ORIGINAL index (int bits);
//                     padded with +inf points to a multiple of 64.
//   level L-1 (leaf)  node j covers spts[8j .. 8j+7]
//   level l < L-1     node j covers nodes 8j .. 8j+7 of level l+1
//   level 0           at most 8 nodes (one group)
// Boxes are stored per GROUP of 8 sibling nodes as 12 float4 (SoA):
//   [lo.x x8][lo.y x8][lo.z x8][hi.x x8][hi.y x8][hi.z x8]
// so that one node visit is 12 coalescable 16-byte loads.  Unused slots hold lo=+inf, hi=-inf,
// whose distance bound is +inf and therefore never qualifies.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxLevels = 8;
constexpr int kBranch = 8;
constexpr int kLeaf = 8;

struct IndexView {
  const float4* spts;
  const float4* box[kMaxLevels];
  int cnt[kMaxLevels];
  int n;
  int nlev;
};

// 6 doubles per point: xx, xy, xz, yy, yz, zz (the 3x3 block of the reference's Matrix4d)
constexpr int kCovStride = 6;

// number of fp64 sums one linearize produces: 21 (upper H) + 6 (b) + 1 (error)
constexpr int kNumSums = 28;

struct RuntimeImpl;

}  // namespace ddlo

// opaque handle bodies ---------------------------------------------------------------------------
struct ddlo_runtime {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
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
  int max_coop_blocks_align = 0;  // co-resident 256-thread blocks of the align kernel
};

struct ddlo_cloud {
  std::atomic<int> refs{1};
  ddlo_runtime* rt = nullptr;
  int n = 0;
  float4* pts = nullptr;  // original order, w = 1
  // index (optional)
  bool has_index = false;
  int npad = 0;           // padded length of spts (multiple of 64)
  float4* spts = nullptr;
  float4* boxes = nullptr;  // all levels, contiguous
  ddlo::IndexView view{};
};

struct ddlo_covs {
  std::atomic<int> refs{1};
  ddlo_runtime* rt = nullptr;
  int n = 0;
  double* c = nullptr;  // n * 6
};
