// Host-side state of one NanoGICP engine and the helpers shared by api.cu, batch.cu and keyframes.cu.
#pragma once

#include "common.cuh"
#include "gicp.cuh"

struct ddlo_gicp {
  ddlo_runtime* rt = nullptr;
  ddlo_params p{};
  ddlo_cloud* src = nullptr;
  ddlo_cloud* tgt = nullptr;
  ddlo_covs* src_cov = nullptr;
  ddlo_covs* tgt_cov = nullptr;
  // per-source-point workspace (correspondences_, sq_distances_, mahalanobis_)
  int ws_n = 0;
  int* corr = nullptr;
  int2* nn_seed = nullptr;
  float* sqd = nullptr;
  double* mahal = nullptr;
  int corr_n = 0;  // number of valid entries (0 after swap/clear/new input: correspondences_.clear())
  double* partials = nullptr;
  int partial_stride = 0;
  ddlo::AlignOut* d_out = nullptr;
  int4* d_dbg = nullptr;  // DDLO_VISIT_STATS builds only
  int d_dbg_n = 0;
  // profiling (off by default, ddlo_gicp_debug_enable): [8][partial_stride][8] block times, then kStampCap + 1 stamps
  unsigned long long* d_prof = nullptr;
  bool profile = false;
  float last_T[16];
  bool has_last_T = false;
  bool align_pending = false;
  int pending_covs_computed = 0;
};

namespace ddlo {

constexpr int kStampCap = 128;

// reference counting (api.cu)
void cloud_set(ddlo_cloud*& slot, ddlo_cloud* c);
void covs_set(ddlo_covs*& slot, ddlo_covs* v);
int cloud_new(ddlo_runtime* rt, int n, ddlo_cloud** out);
int cloud_adopt(ddlo_runtime* rt, float4* pts, int n, ddlo_cloud** out);
int covs_new(ddlo_runtime* rt, int n, ddlo_covs** out);
int use_device(const ddlo_runtime* rt);
int ensure_pinned(ddlo_runtime* rt, size_t bytes);
// enqueue one align on the engine's stream (missing covariances are computed first); nothing is read back
int enqueue_align(ddlo_gicp* g, const float* guess16, int* covs_computed);
// single_launch: the caller will run k_align (the first correspondence search may then be started early, beside the
// covariance kernels); false: the caller only wants the argument record (batched waves)
int prepare_align(ddlo_gicp* g, const float* guess16, int* covs_computed, GicpArgs* a, int* nblocks, bool single_launch = false);
void fill_result(const AlignOut* o, int covs_computed, ddlo_align_result* r);
// batch_align.cu: a wave of problems advanced together (device array of BatchProb records, opaque here)
size_t batch_prob_bytes();
void batch_prob_fill(void* h_probs, int slot, const GicpArgs* args, int nchunks);  // args == nullptr: empty slot
int batch_align_begin(cudaStream_t st, void* d_probs, int n, int* d_active, long long* launches);
int batch_align_round(cudaStream_t st, void* d_probs, int n, int max_chunks, int* d_active, long long* launches);
// the target covariances permuted into the Morton order of `cloud`'s index (cached on the covariance handle)
int ensure_sorted_covs(ddlo_covs* v, ddlo_cloud* cloud, cudaStream_t st);

}  // namespace ddlo
