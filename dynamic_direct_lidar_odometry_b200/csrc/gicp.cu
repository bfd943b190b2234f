// The GICP cost function and the Levenberg-Marquardt / Gauss-Newton driver for ONE registration, on the device.
//
//   k_align                LsqRegistration::computeTransformation, step_lm, step_gn, is_converged
//                                                    (lsq_registration_impl.hpp:96-232)
// built from the device code of gicp_dev.cuh (search_slots, lin_group, err_point, the lm_* state machine).
//
// k_align is ONE cooperative launch per align(), one 1024-thread block per SM; block b owns chunk b of the source
// points (gicp_dev.cuh, Deal).  Every outer iteration a block runs
//   phase A  1-NN of the moved source points in the target octree, one query per PAIR of lanes
//            (knn_pair.cuh), seeded with the previous iteration's match; pairs pull queries from a
//            queue in shared memory and park the matches per slot;
//   phase B  one thread per point: Mahalanobis matrix, residual, the 28 contributions to
//            J^T M J / J^T M e / e^T M e in fp64, summed per group of 32 slots by butterfly;
// then adds the group sums in group order, followed by a grid-wide reduction (per-block partials in L2, grid.sync,
// every block adds them in the same fixed order), the 6x6 LM solve, and the trial-error passes, all without the host.
// Every block evaluates the (tiny) LM controller redundantly from identical sums, which keeps the
// control flow uniform across the grid without a broadcast.  Reductions have a fixed order, so
// results are bit-reproducible run to run, and identical to the batched kernels' (batch_align.cu).
#include <cooperative_groups.h>

#include "gicp_dev.cuh"

namespace cg = cooperative_groups;

namespace ddlo {

constexpr int kAlignWarps = kAlignThreads / 32;
// Source points (slots) of one ROUND of a block: searched from one queue (phase A), then linearized (phase B).  A block
// has 512 lane pairs; a chunk of up to 512 slots (C1, C2, C3: 443) is in flight at once, a larger one (C4: 1 771) refills
// the pairs from the round's queue, so the round should hold the whole chunk if it can: every round ends with the tail of
// its longest queries.
#ifndef DDLO_ALIGN_ROUND
#define DDLO_ALIGN_ROUND 2048
#endif
constexpr int kRoundSlots = DDLO_ALIGN_ROUND;
constexpr int kRoundGroups = kRoundSlots / kGroup;
static_assert(kRoundSlots % kGroup == 0, "a round is whole groups");
#ifndef DDLO_SEARCH_WARPS
#define DDLO_SEARCH_WARPS 32
#endif
constexpr int kSearchWarps = DDLO_SEARCH_WARPS;  // warps that search (16 queries in flight each, refilled from the round's queue)

// dynamic shared memory of the align / step kernels
struct AlignSmem {
  float nn_d[kRoundSlots];  // matches of the current round, parked per slot
  int nn_idx[kRoundSlots];
  int nn_pos[kRoundSlots];
  int next;                 // phase A queue head
  int next_action;          // what the LM controller wants next (kNext*)
  double gs[kRoundGroups][kNumSums];  // group sums of the current round
  double egs[kAlignWarps * 4];        // group sums of an error pass (up to 4096 slots per chunk; larger chunks loop)
  double acc[kNumSums];               // the chunk's sums over the rounds done so far
  double tot[kNumSums];
  unsigned long long t_search;  // %globaltimer when the last warp of this block finished its search (profiling)
  LmShared lm;
};

// linearize over the block's chunk; leaves the chunk's 28 sums in dst[c * stride + blockIdx.x].
// have_prev: nn_seed holds the matches of the previous linearize of the same source cloud.
// matches_ready: the correspondences of this transform are already in corr / nn_seed (k_search_pass0): no search.
__device__ __forceinline__ void linearize_block(const GicpArgs& a, AlignSmem& sm, bool have_prev, double* dst, int stride,
                                                unsigned long long* bt = nullptr, bool matches_ready = false) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kNumSums) sm.acc[threadIdx.x] = 0.0;
  if (threadIdx.x == 0) sm.t_search = 0ull;
  const Deal dl = make_deal(a.ns, gridDim.x, blockIdx.x);
  SearchPark pk{sm.nn_d, sm.nn_idx, sm.nn_pos, &sm.next, sm.lm.Rf, sm.lm.tf, sm.lm.n_lin};
  for (int base = 0; base < dl.nslots; base += kRoundSlots) {
    const int nround = min(kRoundSlots, dl.nslots - base);  // multiple of 16
    if (threadIdx.x == 0) sm.next = 0;
    __syncthreads();  // queue reset; parked matches and group sums of the previous round consumed; acc initialised
    // ---- phase A
    if (warp < kSearchWarps && !matches_ready) {
      search_slots(a, pk, dl, base, nround, have_prev);
      if (bt && lane == 0) atomicMax(&sm.t_search, globaltimer_ns());  // when the block's last warp left the search
    }
    __syncthreads();
    // ---- phase B: one thread per point of the round, whole warps (a group = 32 slots; the last one may be half empty)
    const int ngroups = (nround + kGroup - 1) / kGroup;
    for (int g = warp; g < ngroups; g += kAlignWarps) {
      const unsigned long long t0 = bt ? globaltimer_ns() : 0ull;
      const int slot = g * kGroup + lane;
      int i = -1, j = -1, pos = -1;
      if (slot < nround) {
        i = dl.point(base + slot);
        if (i >= a.ns) {
          i = -1;
        } else if (matches_ready) {
          j = __ldcg(a.corr + i);
          pos = __ldcg(a.nn_seed + i).x;
        } else {
          pos = sm.nn_pos[slot];
          j = store_match(a, i, sm.nn_d[slot], sm.nn_idx[slot], pos);
        }
      }
      const double v = lin_group(a, sm.lm.x0, i, j, pos);
      if (lane < kNumSums) sm.gs[g][lane] = v;
      if (bt && lane == 0) {
        const unsigned long long dt = globaltimer_ns() - t0;
        atomicMax(bt + 4, dt);
        atomicMax(bt + 5, 0xffffffffull - dt);
      }
    }
    __syncthreads();
    // ---- the round's group sums, in group order
    if (threadIdx.x < kNumSums) {
      double v = sm.acc[threadIdx.x];
      for (int g = 0; g < ngroups; ++g) v += sm.gs[g][threadIdx.x];
      sm.acc[threadIdx.x] = v;
    }
  }
  __syncthreads();
  if (bt && threadIdx.x == 0) bt[6] = globaltimer_ns();
  if (threadIdx.x < kNumSums) __stcg(dst + (size_t)threadIdx.x * stride + blockIdx.x, sm.acc[threadIdx.x]);
  __syncthreads();
  if (bt && threadIdx.x == 0) bt[7] = globaltimer_ns();
}

// compute_error over the block's chunk; the chunk's sum (group sums added in group order) goes to dst[blockIdx.x]
__device__ __forceinline__ void error_block(const GicpArgs& a, const Iso3& T, AlignSmem& sm, double* dst) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Deal dl = make_deal(a.ns, gridDim.x, blockIdx.x);
  constexpr int kSpan = kAlignWarps * 4;  // groups per sweep of the block
  double total = 0.0;                     // thread 0 only
  const int ngroups = (dl.nslots + kGroup - 1) / kGroup;
  for (int g0 = 0; g0 < ngroups; g0 += kSpan) {
    for (int g = g0 + warp; g < min(ngroups, g0 + kSpan); g += kAlignWarps) {
      const int slot = g * kGroup + lane;
      int i = -1;
      if (slot < dl.nslots) {
        i = dl.point(slot);
        if (i >= a.ns) i = -1;
      }
      const double e = err_group(a, T, i);
      if (lane == 0) sm.egs[g - g0] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int g = g0; g < min(ngroups, g0 + kSpan); ++g) total += sm.egs[g - g0];
    __syncthreads();
  }
  if (threadIdx.x == 0) __stcg(dst + blockIdx.x, total);
  __syncthreads();
}

// phase tags of the timeline block 0 leaves in GicpArgs::stamps (profiling runs only)
enum { kTagSearchDone = 10, kTagStart = 1, kTagLinDone = 2, kTagLinSynced = 3, kTagLinSummed = 4, kTagSolved = 5, kTagErrDone = 6, kTagErrSynced = 7, kTagDecided = 8, kTagEnd = 9 };
#define DDLO_STAMP(tag)                                                                                         \
  do {                                                                                                          \
    if (a.stamps && blockIdx.x == 0 && threadIdx.x == 0 && n_stamps < 128)                                      \
      a.stamps[1 + n_stamps++] = ((unsigned long long)(tag) << 56) | (globaltimer_ns() & 0x00ffffffffffffffull); \
  } while (0)

__global__ void __launch_bounds__(kAlignThreads, kAlignBlocksPerSM) k_align(const GicpArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AlignSmem& sm = *reinterpret_cast<AlignSmem*>(smem_raw);
  LmShared& s = sm.lm;
  const int nblk = gridDim.x;
  int seq = 0;  // reduction counter: partial buffers alternate so a fast block can not overwrite
                // sums a slow block is still reading
  int n_stamps = 0;
  DDLO_STAMP(kTagStart);

  if (threadIdx.x == 0) sm.next_action = lm_start(s, a);
  __syncthreads();

  while (sm.next_action != kNextDone) {
    double* part = a.partials + (size_t)(seq & 1) * kNumSums * a.partial_stride;
    ++seq;
    if (sm.next_action == kNextLinearize) {
      // ---- linearize(x0) ---------------------------------------------------------------------
      const int pass = s.n_lin;
      unsigned long long* bt = (a.blk_times && pass < 8) ? a.blk_times + ((size_t)pass * nblk + blockIdx.x) * 8 : nullptr;
      if (bt && threadIdx.x == 0) bt[0] = globaltimer_ns();
      linearize_block(a, sm, pass > 0, part, a.partial_stride, bt, pass == 0 && a.pass0_done != 0);
      if (bt && threadIdx.x == 0) {
        bt[1] = sm.t_search;
        bt[2] = globaltimer_ns();
      }
      if (a.stamps && blockIdx.x == 0 && threadIdx.x == 0 && n_stamps < 128)
        a.stamps[1 + n_stamps++] = ((unsigned long long)kTagSearchDone << 56) | (sm.t_search & 0x00ffffffffffffffull);
      DDLO_STAMP(kTagLinDone);
      grid.sync();
      if (bt && threadIdx.x == 0) bt[3] = globaltimer_ns();
      DDLO_STAMP(kTagLinSynced);
      grid_sum<kNumSums>(part, a.partial_stride, nblk, sm.tot);
      DDLO_STAMP(kTagLinSummed);
      if (threadIdx.x == 0) sm.next_action = lm_on_linearized(s, a, sm.tot);
      __syncthreads();
      DDLO_STAMP(kTagSolved);
    } else {
      // ---- compute_error(xi): one trial of step_lm ---------------------------------------------
      error_block(a, s.xi, sm, part);
      DDLO_STAMP(kTagErrDone);
      grid.sync();
      DDLO_STAMP(kTagErrSynced);
      grid_sum<1>(part, a.partial_stride, nblk, sm.tot);
      if (threadIdx.x == 0) sm.next_action = lm_on_error(s, a, sm.tot[0]);
      __syncthreads();
      DDLO_STAMP(kTagDecided);
    }
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    write_align_out(a, s, a.out);
    DDLO_STAMP(kTagEnd);
    if (a.stamps) a.stamps[0] = (unsigned long long)n_stamps;
  }
}

// update_correspondences at the guess, for all chunks of the registration, as a kernel of its own (256-thread blocks
// that share an SM with the covariance kernels; the chunks are k_align's, the matches land where k_align's first
// linearize would have put them)
constexpr int kP0Threads = 256;
constexpr int kP0Round = 1024;
__global__ void __launch_bounds__(kP0Threads, 4) k_search_pass0(const GicpArgs a) {
  __shared__ float s_d[kP0Round];
  __shared__ int s_idx[kP0Round], s_pos[kP0Round];
  __shared__ int s_next;
  __shared__ float s_T[12];
  if (threadIdx.x == 0) {
    Iso3 x0;
    iso_from_colmajor(a.guess, x0);
    iso_to_float(x0, s_T, s_T + 9);
  }
  const Deal dl = make_deal(a.ns, gridDim.x, blockIdx.x);
  SearchPark pk{s_d, s_idx, s_pos, &s_next, s_T, s_T + 9, 0};
  for (int base = 0; base < dl.nslots; base += kP0Round) {
    const int nround = min(kP0Round, dl.nslots - base);
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    search_slots(a, pk, dl, base, nround, false);
    __syncthreads();
    for (int t = threadIdx.x; t < nround; t += kP0Threads) {
      const int i = dl.point(base + t);
      if (i < a.ns) store_match(a, i, s_d[t], s_idx[t], s_pos[t]);
    }
    __syncthreads();
  }
}

// ---- stepwise hooks (parity tests compare H, b, error, correspondences with the oracle) ---------
__global__ void __launch_bounds__(kAlignThreads, kAlignBlocksPerSM) k_linearize_step(const GicpArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AlignSmem& sm = *reinterpret_cast<AlignSmem*>(smem_raw);
  if (threadIdx.x == 0) {
    iso_from_colmajor(a.T_step, sm.lm.x0);
    iso_to_float(sm.lm.x0, sm.lm.Rf, sm.lm.tf);
    sm.lm.n_lin = 0;
  }
  __syncthreads();
  linearize_block(a, sm, false, a.partials, a.partial_stride);
}

__global__ void __launch_bounds__(kAlignThreads, kAlignBlocksPerSM) k_error_step(const GicpArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AlignSmem& sm = *reinterpret_cast<AlignSmem*>(smem_raw);
  if (threadIdx.x == 0) iso_from_colmajor(a.T_step, sm.lm.xi);
  __syncthreads();
  error_block(a, sm.lm.xi, sm, a.partials);
}

__global__ void __launch_bounds__(256) k_sum_partials(const double* partials, int stride, int nblk, int ncomp, AlignOut* out) {
  __shared__ double s_tot[kNumSums];
  if (ncomp == 1)
    grid_sum<1>(partials, stride, nblk, s_tot);
  else
    grid_sum<kNumSums>(partials, stride, nblk, s_tot);
  if (threadIdx.x < ncomp) out->sums[threadIdx.x] = s_tot[threadIdx.x];
}

// getResiduals(std::vector<Eigen::Vector3f>&, trans) (nano_gicp_impl.hpp:199-222): float B - trans*A
__global__ void __launch_bounds__(256) k_residual_vectors(const float4* __restrict__ src, const float4* __restrict__ tgt,
                                                          const int* __restrict__ corr, int n, const float* __restrict__ T /*Rf9,tf3*/,
                                                          float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = corr[i];
  float rx = 0.0f, ry = 0.0f, rz = 0.0f;
  if (j >= 0) {
    const float4 p = src[i], b = tgt[j];
    rx = __fsub_rn(b.x, xform_f(T + 0, T[9], p.x, p.y, p.z));
    ry = __fsub_rn(b.y, xform_f(T + 3, T[10], p.x, p.y, p.z));
    rz = __fsub_rn(b.z, xform_f(T + 6, T[11], p.x, p.y, p.z));
  }
  out[3 * i + 0] = rx;
  out[3 * i + 1] = ry;
  out[3 * i + 2] = rz;
}

// pcl::transformPointCloud (float)
__global__ void __launch_bounds__(256) k_transform_cloud(const float4* __restrict__ src, int n, const float* __restrict__ T,
                                                         float4* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = src[i];
  dst[i] = make_float4(xform_f(T + 0, T[9], p.x, p.y, p.z), xform_f(T + 3, T[10], p.x, p.y, p.z), xform_f(T + 6, T[11], p.x, p.y, p.z), 1.0f);
}

// ---- launchers ----------------------------------------------------------------------------------
// function attributes are per device: remember which devices have them (a process may drive several GPUs)
static int set_smem_attrs() {
  static std::atomic<unsigned long long> done_mask{0};
  int dev = 0;
  DDLO_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (done_mask.load() & bit) return DDLO_OK;
  const int bytes = (int)sizeof(AlignSmem);
  DDLO_CUDA(cudaFuncSetAttribute(k_align, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  DDLO_CUDA(cudaFuncSetAttribute(k_linearize_step, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  DDLO_CUDA(cudaFuncSetAttribute(k_error_step, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done_mask.fetch_or(bit);
  return DDLO_OK;
}

int gicp_max_coop_blocks(int device, int* blocks_per_sm) {
  DDLO_TRY(set_smem_attrs());
  int per_sm = 0;
  DDLO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_align, kAlignThreads, sizeof(AlignSmem)));
  int coop = 0;
  DDLO_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  if (!coop) return fail(DDLO_E_UNSUPPORTED, "device lacks cooperative launch");
  *blocks_per_sm = per_sm;
  return DDLO_OK;
}

int gicp_blocks_for(int ns, int max_blocks) {
  const int want = (ns + 15) / 16;  // at least one group of 16 points per block
  return std::max(1, std::min(want, max_blocks));
}

int launch_align(ddlo_runtime* rt, const GicpArgs& args, int blocks) {
  DDLO_TRY(set_smem_attrs());
  void* kargs[] = {const_cast<GicpArgs*>(&args)};
  DDLO_CUDA(cudaLaunchCooperativeKernel((const void*)k_align, dim3(blocks), dim3(kAlignThreads), kargs, sizeof(AlignSmem), rt->stream));
  rt->launches += 1;
  return DDLO_OK;
}

int launch_search_pass0(ddlo_runtime* rt, cudaStream_t st, const GicpArgs& args, int chunks) {
  k_search_pass0<<<chunks, kP0Threads, 0, st>>>(args);
  rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

int launch_linearize_step(ddlo_runtime* rt, const GicpArgs& args, int blocks) {
  DDLO_TRY(set_smem_attrs());
  k_linearize_step<<<blocks, kAlignThreads, sizeof(AlignSmem), rt->stream>>>(args);
  k_sum_partials<<<1, 256, 0, rt->stream>>>(args.partials, args.partial_stride, blocks, kNumSums, args.out);
  rt->launches += 2;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

int launch_error_step(ddlo_runtime* rt, const GicpArgs& args, int blocks) {
  DDLO_TRY(set_smem_attrs());
  k_error_step<<<blocks, kAlignThreads, sizeof(AlignSmem), rt->stream>>>(args);
  k_sum_partials<<<1, 256, 0, rt->stream>>>(args.partials, args.partial_stride, blocks, 1, args.out);
  rt->launches += 2;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

static void pack_T(const float* T16, float* out12) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) out12[3 * i + j] = T16[4 * j + i];
    out12[9 + i] = T16[12 + i];
  }
}

int launch_residual_vectors(ddlo_runtime* rt, const float4* src, const float4* tgt, const int* corr, int n, const float* T16_host,
                            float* d_out3) {
  float h[12];
  pack_T(T16_host, h);
  float* dT = reinterpret_cast<float*>(rt->d_scratch);
  DDLO_CUDA(cudaMemcpyAsync(dT, h, sizeof(h), cudaMemcpyHostToDevice, rt->stream));
  k_residual_vectors<<<(n + 255) / 256, 256, 0, rt->stream>>>(src, tgt, corr, n, dT, d_out3);
  rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

int launch_transform_cloud(ddlo_runtime* rt, const float4* src, int n, const float* T16_host, float4* dst) {
  float h[12];
  pack_T(T16_host, h);
  float* dT = reinterpret_cast<float*>(rt->d_scratch);
  DDLO_CUDA(cudaMemcpyAsync(dT, h, sizeof(h), cudaMemcpyHostToDevice, rt->stream));
  k_transform_cloud<<<(n + 255) / 256, 256, 0, rt->stream>>>(src, n, dT, dst);
  rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

}  // namespace ddlo
