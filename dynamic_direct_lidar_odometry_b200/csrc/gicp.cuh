// Device-side data of one NanoGICP engine and the launchers implemented in gicp.cu.
#pragma once

#include "common.cuh"
#include "math.cuh"

namespace ddlo {

// what align() leaves behind for the host (one small D2H copy)
struct AlignOut {
  float final_transformation[16];  // column-major
  double final_hessian[36];        // column-major (symmetric)
  int flags;
  int nr_iterations;
  int n_linearize;
  int n_compute_error;
  double final_error;
  double lm_lambda;
  double sums[kNumSums];  // stepwise hooks: H upper / b / error of the last reduction
};

struct GicpArgs {
  IndexView tgt;            // target kNN index (Morton-ordered points + box tree)
  const float4* src_pts;    // ns, original order
  const float* src_lattice; // lattice record of the source index (non-finite counter at [4]) or nullptr
  const double* src_cov;    // ns * 6
  const float4* tgt_pts;    // nt, original order (compute_error gathers matched points here)
  const double* tgt_cov;    // nt * 6, MORTON order of the target index (gathers of neighbouring matches share DRAM pages)
  int ns;
  int* corr;                // correspondences_   (ns)
  int2* nn_seed;            // per source point {match position in spts, node above its leaf} or {-1,-1}: seeds the next search
  float* sqd;               // sq_distances_      (ns)
  double* mahal;            // mahalanobis_       (ns * 6)
  double* partials;         // [2][kNumSums][partial_stride]
  int partial_stride;
  int max_iterations;
  int optimizer;
  int lm_max_iterations;
  double thr2;              // corr_dist_threshold_^2
  double trans_eps, rot_eps, lm_init_lambda_factor;
  int pass0_done;           // the first linearize's correspondences are already in corr / nn_seed / sqd (k_search_pass0)
  float guess[16];          // column-major Eigen::Matrix4f
  double T_step[16];        // stepwise hooks: transform to evaluate at (column-major)
  AlignOut* out;
  int4* dbg_visits;         // DDLO_VISIT_STATS builds only: [pass < 4][ns] {node visits, leaf scans, warp steps, 0}
  // profiling, both null unless ddlo_gicp_debug_enable was called:
  unsigned long long* blk_times;  // [pass < 8][block][8] %globaltimer at pass start / search done / phase B done / after grid sync, then max and (2^32-1 - min) lin_point duration
  unsigned long long* stamps;     // [0] = count, then the phase timeline of block 0: tag << 56 | %globaltimer ns
};

int gicp_max_coop_blocks(int device, int* blocks_per_sm);
int launch_align(ddlo_runtime* rt, const GicpArgs& args, int blocks);
// update_correspondences of the FIRST linearize (transform = the guess) as a kernel of its own on `st`: lets the
// search run beside the covariance kernels when align() has to compute covariances first
int launch_search_pass0(ddlo_runtime* rt, cudaStream_t st, const GicpArgs& args, int chunks);
int launch_linearize_step(ddlo_runtime* rt, const GicpArgs& args, int blocks);
int launch_error_step(ddlo_runtime* rt, const GicpArgs& args, int blocks);
int launch_residual_vectors(ddlo_runtime* rt, const float4* src, const float4* tgt, const int* corr, int n, const float* T16_host,
                            float* d_out3);
int launch_transform_cloud(ddlo_runtime* rt, const float4* src, int n, const float* T16_host, float4* dst);

constexpr int kAlignThreads = 1024;  // one block per SM: 128 search sub-warps (512 threads with 128 registers was tried: slower search)
constexpr int kAlignBlocksPerSM = 1;  // (768 x 2 was tried: more warps, but the fp64 phase spills and the pass gets slower)
int gicp_blocks_for(int ns, int max_blocks);

}  // namespace ddlo
