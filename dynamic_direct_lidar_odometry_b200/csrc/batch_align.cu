// LM / GN registration of MANY independent problems per launch (the align stage of the batched path, batch.cu).
//
// k_align (gicp.cu) runs one registration as one cooperative launch and is built for latency: one block per SM, grid
// barriers between the phases, every SM waiting for the slowest query of the pass.  For a batch that shape wastes the
// machine, so here the same device code (gicp_dev.cuh) is driven the other way round: a ROUND is three ordinary
// launches over all problems of a wave,
//     k_batch_search   1-NN of every source point of every problem that is due for a linearize
//     k_batch_lin      Mahalanobis + H, b, error contributions of those problems; the last block of a problem to
//                      finish adds the chunk sums and advances the problem's LM state machine
//     k_batch_err      compute_error of every problem that is due for an LM trial; last block advances the state
// with 256-thread blocks, one per (problem, chunk), scheduled freely over the SMs: nothing waits at a barrier wider
// than a block, and problems at different iterations share the launches.  The chunks, groups and summation orders are
// exactly k_align's, so a problem's result is bit-identical to what ddlo_gicp_align returns for it.
#include <algorithm>

#include "engine.cuh"
#include "gicp_dev.cuh"

namespace ddlo {

// threads of a search block (one block per problem and chunk).  Measured on the C2 batch, registrations/s: 512 threads
// x 2 blocks per SM 4 490, 256 x 4 4 810, 128 x 8 4 945, 64 x 16 4 900 - smaller blocks leave fewer pairs idle while the
// chunk's last long queries finish.  Serving 1024 or 2048 CONSECUTIVE points per block instead of a chunk (the search
// does not depend on how the sums are chunked) was measured too: 4 890 / 4 620, no gain.
#ifndef DDLO_BATCH_SEARCH_THREADS
#define DDLO_BATCH_SEARCH_THREADS 128
#endif
constexpr int kBThreads = DDLO_BATCH_SEARCH_THREADS;
constexpr int kBRound = 1024;  // slots of a chunk searched per queue round (parking space in shared memory)
// threads of the linearize / error kernels.  Measured (registrations/s on the C2 batch, one box per line):
// 512 threads 4 500 against 4 640 with 256; 256 / 128 / 64 threads 4 996 / 5 051 / 5 097.  The sums do not depend on it
// (group sums are added in group order whatever the number of warps).
#ifndef DDLO_BATCH_LIN_THREADS
#define DDLO_BATCH_LIN_THREADS 64
#endif
constexpr int kLThreads = DDLO_BATCH_LIN_THREADS;
constexpr int kLWarps = kLThreads / 32;
#ifndef DDLO_BATCH_BLOCKS_PER_SM
#define DDLO_BATCH_BLOCKS_PER_SM (1024 / DDLO_BATCH_SEARCH_THREADS)
#endif
constexpr int kBBlocksPerSM = DDLO_BATCH_BLOCKS_PER_SM;  // 1024 threads per SM: 64 registers per thread, as in k_align

struct BatchProb {
  GicpArgs a;
  LmShared lm;
  int next;     // kNext*
  int nchunks;  // blocks a single align of this problem uses = chunks of its source points; 0: empty slot
  unsigned ticket;
  int pad_;
};

__global__ void __launch_bounds__(128) k_batch_begin(BatchProb* probs, int n, int* n_active) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  BatchProb& P = probs[p];
  P.ticket = 0u;
  if (P.nchunks <= 0) {
    P.next = kNextDone;
    return;
  }
  LmShared s;
  const int next = lm_start(s, P.a);
  P.lm = s;
  P.next = next;
  if (next == kNextDone)
    write_align_out(P.a, s, P.a.out);
  else
    atomicAdd(n_active, 1);
}

// the problem's argument record, copied into shared memory once per block (the kernels of gicp.cu get theirs as a
// kernel parameter, i.e. from the constant bank)
__device__ __forceinline__ void load_args(const BatchProb& P, GicpArgs& s_a) {
  static_assert(sizeof(GicpArgs) % sizeof(int) == 0, "GicpArgs is copied word by word");
  const int* src = reinterpret_cast<const int*>(&P.a);
  int* dst = reinterpret_cast<int*>(&s_a);
  for (int w = threadIdx.x; w < (int)(sizeof(GicpArgs) / sizeof(int)); w += blockDim.x) dst[w] = __ldg(src + w);
}

__global__ void __launch_bounds__(kBThreads, kBBlocksPerSM) k_batch_search(BatchProb* probs) {
  BatchProb& P = probs[blockIdx.y];
  if ((int)blockIdx.x >= P.nchunks || P.next != kNextLinearize) return;
  __shared__ float s_d[kBRound];
  __shared__ int s_idx[kBRound], s_pos[kBRound];
  __shared__ int s_next;
  __shared__ float s_T[12];
  __shared__ GicpArgs s_a;
  load_args(P, s_a);
  if (threadIdx.x < 9) s_T[threadIdx.x] = P.lm.Rf[threadIdx.x];
  if (threadIdx.x < 3) s_T[9 + threadIdx.x] = P.lm.tf[threadIdx.x];
  __syncthreads();
  const GicpArgs& a = s_a;
  const bool have_prev = P.lm.n_lin > 0;
  const Deal dl = make_deal(a.ns, P.nchunks, blockIdx.x);
  SearchPark pk{s_d, s_idx, s_pos, &s_next, s_T, s_T + 9, P.lm.n_lin};
  for (int base = 0; base < dl.nslots; base += kBRound) {
    const int nround = min(kBRound, dl.nslots - base);
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    search_slots(a, pk, dl, base, nround, have_prev);
    __syncthreads();
    for (int t = threadIdx.x; t < nround; t += kBThreads) {
      const int i = dl.point(base + t);
      if (i < a.ns) store_match(a, i, s_d[t], s_idx[t], s_pos[t]);
    }
    __syncthreads();
  }
}

// the last block of a problem to arrive here sums the chunk partials and runs the LM controller
template <int NCOMP>
__device__ __forceinline__ void finish_problem(BatchProb& P, const GicpArgs& a, int* n_active, double* s_tot, LmShared& s_lm, int* s_last) {
  __threadfence();  // this block's partial sums are visible before its ticket
  __syncthreads();
  if (threadIdx.x == 0) *s_last = atomicAdd(&P.ticket, 1u) == (unsigned)(P.nchunks - 1) ? 1 : 0;
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  grid_sum<NCOMP>(a.partials, a.partial_stride, P.nchunks, s_tot);
  // the controller works on a copy of the problem's state in shared memory
  {
    const int words = (int)(sizeof(LmShared) / sizeof(int));
    const int* src = reinterpret_cast<const int*>(&P.lm);
    int* dst = reinterpret_cast<int*>(&s_lm);
    for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = __ldcg(src + w);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int next = NCOMP == 1 ? lm_on_error(s_lm, a, s_tot[0]) : lm_on_linearized(s_lm, a, s_tot);
    if (next == kNextDone) {
      write_align_out(a, s_lm, a.out);
      atomicSub(n_active, 1);
    }
    *s_last = next;
  }
  __syncthreads();
  {
    const int words = (int)(sizeof(LmShared) / sizeof(int));
    const int* src = reinterpret_cast<const int*>(&s_lm);
    int* dst = reinterpret_cast<int*>(&P.lm);
    for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  }
  if (threadIdx.x == 0) {
    P.ticket = 0u;
    P.next = *s_last;
  }
}

__global__ void __launch_bounds__(kLThreads, 1024 / kLThreads) k_batch_lin(BatchProb* probs, int* n_active) {
  BatchProb& P = probs[blockIdx.y];
  if ((int)blockIdx.x >= P.nchunks || P.next != kNextLinearize) return;
  __shared__ double s_gs[kLWarps][kNumSums];
  __shared__ double s_tot[kNumSums];
  __shared__ Iso3 s_x0;
  __shared__ LmShared s_lm;
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ GicpArgs s_a;
  load_args(P, s_a);
  if (threadIdx.x < 9) s_x0.r[threadIdx.x] = P.lm.x0.r[threadIdx.x];
  if (threadIdx.x < 3) s_x0.t[threadIdx.x] = P.lm.x0.t[threadIdx.x];
  __syncthreads();
  const GicpArgs& a = s_a;
  const Deal dl = make_deal(a.ns, P.nchunks, blockIdx.x);
  double acc = 0.0;  // threads < 28: the chunk's sum of component threadIdx.x
  for (int base = 0; base < dl.nslots; base += kLThreads) {
    const int slot = base + threadIdx.x;
    const int ngroups = min(kLWarps, (dl.nslots - base + kGroup - 1) / kGroup);
    if (warp < ngroups) {
      int i = -1, j = -1, pos = -1;
      if (slot < dl.nslots) {
        i = dl.point(slot);
        if (i < a.ns) {
          j = __ldcg(a.corr + i);
          pos = __ldcg(a.nn_seed + i).x;
        } else {
          i = -1;
        }
      }
      const double v = lin_group(a, s_x0, i, j, pos);
      if (lane < kNumSums) s_gs[warp][lane] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNumSums)
      for (int g = 0; g < ngroups; ++g) acc += s_gs[g][threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x < kNumSums) __stcg(a.partials + (size_t)threadIdx.x * a.partial_stride + blockIdx.x, acc);
  finish_problem<kNumSums>(P, a, n_active, s_tot, s_lm, &s_last);
}

__global__ void __launch_bounds__(kLThreads, 1024 / kLThreads) k_batch_err(BatchProb* probs, int* n_active) {
  BatchProb& P = probs[blockIdx.y];
  if ((int)blockIdx.x >= P.nchunks || P.next != kNextError) return;
  constexpr int kSpan = 128;  // groups per sweep (k_align's error pass sweeps 128 groups too; the order is the group order anyway)
  __shared__ double s_egs[kSpan];
  __shared__ double s_tot[kNumSums];
  __shared__ Iso3 s_xi;
  __shared__ LmShared s_lm;
  __shared__ int s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ GicpArgs s_a;
  load_args(P, s_a);
  if (threadIdx.x < 9) s_xi.r[threadIdx.x] = P.lm.xi.r[threadIdx.x];
  if (threadIdx.x < 3) s_xi.t[threadIdx.x] = P.lm.xi.t[threadIdx.x];
  __syncthreads();
  const GicpArgs& a = s_a;
  const Deal dl = make_deal(a.ns, P.nchunks, blockIdx.x);
  const int ngroups = (dl.nslots + kGroup - 1) / kGroup;
  double total = 0.0;  // thread 0 only
  for (int g0 = 0; g0 < ngroups; g0 += kSpan) {
    for (int g = g0 + warp; g < min(ngroups, g0 + kSpan); g += kLWarps) {
      const int slot = g * kGroup + lane;
      int i = -1;
      if (slot < dl.nslots) {
        i = dl.point(slot);
        if (i >= a.ns) i = -1;
      }
      const double e = err_group(a, s_xi, i);
      if (lane == 0) s_egs[g - g0] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int g = g0; g < min(ngroups, g0 + kSpan); ++g) total += s_egs[g - g0];
    __syncthreads();
  }
  if (threadIdx.x == 0) __stcg(a.partials + blockIdx.x, total);
  finish_problem<1>(P, a, n_active, s_tot, s_lm, &s_last);
}

// ---- host side ------------------------------------------------------------------------------------------------------
size_t batch_prob_bytes() { return sizeof(BatchProb); }

void batch_prob_fill(void* h_probs, int slot, const GicpArgs* args, int nchunks) {
  BatchProb& P = static_cast<BatchProb*>(h_probs)[slot];
  if (args) P.a = *args;
  P.nchunks = args ? nchunks : 0;
  P.next = kNextDone;
  P.ticket = 0u;
}

int batch_align_begin(cudaStream_t st, void* d_probs, int n, int* d_active, long long* launches) {
  DDLO_CUDA(cudaMemsetAsync(d_active, 0, sizeof(int), st));
  k_batch_begin<<<(n + 127) / 128, 128, 0, st>>>(static_cast<BatchProb*>(d_probs), n, d_active);
  *launches += 1;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

// one round: every problem of the wave advances by one linearize and/or one LM trial
int batch_align_round(cudaStream_t st, void* d_probs, int n, int max_chunks, int* d_active, long long* launches) {
  const dim3 grid((unsigned)max_chunks, (unsigned)n);
  BatchProb* P = static_cast<BatchProb*>(d_probs);
  k_batch_search<<<grid, kBThreads, 0, st>>>(P);
  k_batch_lin<<<grid, kLThreads, 0, st>>>(P, d_active);
  k_batch_err<<<grid, kLThreads, 0, st>>>(P, d_active);
  *launches += 3;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

}  // namespace ddlo
