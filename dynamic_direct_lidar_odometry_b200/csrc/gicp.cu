// The GICP cost function and the Levenberg-Marquardt / Gauss-Newton driver, on the device.
//
//   nn_phase + lin_point   NanoGICP::update_correspondences + the body of linearize
//                                                    (nano_gicp_impl.hpp:235-275, 292-328)
//   err_point              the body of compute_error (nano_gicp_impl.hpp:349-368)
//   k_align                LsqRegistration::computeTransformation, step_lm, step_gn, is_converged
//                                                    (lsq_registration_impl.hpp:96-232)
//
// k_align is ONE cooperative launch per align(), one 1024-thread block per SM.  The source points
// are dealt to the blocks 16 at a time.  Every outer iteration a block runs
//   phase A  1-NN of the moved source points in the target octree, one query per PAIR of lanes
//            (knn_pair.cuh), seeded with the previous iteration's match; pairs pull queries from a
//            queue in shared memory and park the matches per slot;
//   phase B  one thread per point: Mahalanobis matrix, residual, the 28 contributions to
//            J^T M J / J^T M e / e^T M e in fp64, parked per slot in shared memory;
// then adds the parked contributions in a fixed order, followed by a grid-wide reduction (per-block partials in L2, grid.sync, every block adds them in the
// same fixed order), the 6x6 LM solve, and the trial-error passes, all without the host.  Every
// block evaluates the (tiny) LM controller redundantly from identical sums, which keeps the
// control flow uniform across the grid without a broadcast.  Reductions have a fixed order, so
// results are bit-reproducible run to run.
#include <cooperative_groups.h>

#include "gicp.cuh"
#include "knn.cuh"
#include "knn_pair.cuh"

namespace cg = cooperative_groups;

namespace ddlo {

constexpr int kAlignWarps = kAlignThreads / 32;
constexpr int kAlignPairs = kAlignThreads / 2;  // source points (slots) of one round of a block
#ifndef DDLO_SEARCH_WARPS
#define DDLO_SEARCH_WARPS 32
#endif
constexpr int kSearchWarps = DDLO_SEARCH_WARPS;  // warps that search (16 queries in flight each, refilled from the round's queue)

// sum layout: [0..5] H_rr upper, [6..14] H_rt row-major, [15..20] H_tt upper, [21..23] b_r,
// [24..26] b_t, [27] sum of e^T M e
struct LmShared {
  Iso3 x0, xi, delta;
  float Rf[9], tf[3];  // float cast of the transform used for the 1-NN queries (:240)
  double H[36], b[6], d[6];
  double y0, yi, lambda, nu, final_error;
  double final_H[36];
  int action, converged, step_ok, lm_failed, n_lin, n_err, nr_iter;
};

// dynamic shared memory of the align / step kernels
struct AlignSmem {
  double contrib[kNumSums * kAlignPairs];  // [component][slot]: per-point contributions of the current round
  float nn_d[kAlignPairs];                 // matches of the current round, parked per slot
  int nn_idx[kAlignPairs];
  int nn_pos[kAlignPairs];
  int next;                                // phase A queue head
  double red[kAlignWarps];
  double acc[kNumSums];  // the block's sums over the rounds done so far
  double tot[kNumSums];
  unsigned long long t_search;  // %globaltimer when the last warp of this block finished its search (profiling)
  LmShared lm;
};

__device__ __forceinline__ void iso_to_float(const Iso3& T, float* Rf, float* tf) {
  for (int i = 0; i < 9; ++i) Rf[i] = (float)T.r[i];
  for (int i = 0; i < 3; ++i) tf[i] = (float)T.t[i];
}

// Eigen evaluates Transform * Vector4 coefficient-wise with a pairwise unrolled sum:
// (r0*x + r1*y) + (r2*z + t*1).  In float this order is observable in the 1-NN query, so it is
// spelled out with non-contracting intrinsics.
__device__ __forceinline__ float xform_f(const float* r, float t, float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(r[0], x), __fmul_rn(r[1], y)), __fadd_rn(__fmul_rn(r[2], z), t));
}
__device__ __forceinline__ double xform_d(const double* r, double t, double x, double y, double z) {
  return (r[0] * x + r[1] * y) + (r[2] * z + t);
}

__device__ __forceinline__ Sym3 load_sym3(const double* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  return Sym3{a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ Sym3 load_sym3_cg(const double* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2);
  return Sym3{a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ void store_sym3(double* p, const Sym3& s) {
  double2* q = reinterpret_cast<double2*>(p);
  q[0] = make_double2(s.xx, s.xy);
  q[1] = make_double2(s.xz, s.yy);
  q[2] = make_double2(s.yz, s.zz);
}

__device__ __forceinline__ double quad_form(const Sym3& M, double ex, double ey, double ez, double& mx, double& my, double& mz) {
  mx = M.xx * ex + M.xy * ey + M.xz * ez;
  my = M.xy * ex + M.yy * ey + M.yz * ez;
  mz = M.xz * ex + M.yz * ey + M.zz * ez;
  return ex * mx + ey * my + ez * mz;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");
  return t;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Phase B for one point (one thread): the Mahalanobis matrix and the point's 28 contributions,
// written to contrib[component * kAlignPairs].  Every slot of the round gets its column written
// (zeros without a correspondence), so the block sum needs no mask.
// Kept out of line so that its fp64 register appetite does not leak into the search loop.
__device__ __noinline__ void lin_point(const GicpArgs& a, const LmShared& s, bool valid, int i, float nn_d, int nn_idx, int nn_pos,
                                       double* __restrict__ contrib) {
  int j = -1;
  if (valid) {
    a.sqd[i] = nn_d;
    const bool found = nn_idx != kIdxSentinel;
    j = (found && (double)nn_d < a.thr2) ? nn_idx : -1;
    a.corr[i] = j;
    // seed of the next search (kept even beyond the distance threshold): where the match sits in the
    // Morton order and the node above its leaf
    a.nn_seed[i] = found ? make_int2(nn_pos, __ldg(a.tgt.node_of_point + nn_idx)) : make_int2(-1, -1);
  }
  if (j < 0) {
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) contrib[k * kAlignPairs] = 0.0;
    return;
  }
  const float4 pa = __ldg(a.src_pts + i);
  const float4 pb = __ldg(a.tgt.spts + nn_pos);
  const Sym3 CA = load_sym3(a.src_cov + (size_t)i * kCovStride);
  const Sym3 CB = load_sym3(a.tgt_cov + (size_t)nn_pos * kCovStride);
  Sym3 RCR = sym3_rotate(s.x0.r, CA);
  RCR.xx += CB.xx, RCR.xy += CB.xy, RCR.xz += CB.xz, RCR.yy += CB.yy, RCR.yz += CB.yz, RCR.zz += CB.zz;
  const Sym3 M = sym3_inverse(RCR);
  store_sym3(a.mahal + (size_t)i * kCovStride, M);

  const double x = xform_d(s.x0.r + 0, s.x0.t[0], (double)pa.x, (double)pa.y, (double)pa.z);
  const double y = xform_d(s.x0.r + 3, s.x0.t[1], (double)pa.x, (double)pa.y, (double)pa.z);
  const double z = xform_d(s.x0.r + 6, s.x0.t[2], (double)pa.x, (double)pa.y, (double)pa.z);
  const double ex = (double)pb.x - x, ey = (double)pb.y - y, ez = (double)pb.z - z;
  double mex, mey, mez;
#define DDLO_C(k) contrib[(k) * kAlignPairs]
  DDLO_C(27) = quad_form(M, ex, ey, ez, mex, mey, mez);
  // G = S^T M with S = skew(T p_A);  J = [S | -I]
  const double g00 = z * M.xy - y * M.xz, g01 = z * M.yy - y * M.yz, g02 = z * M.yz - y * M.zz;
  const double g10 = x * M.xz - z * M.xx, g11 = x * M.yz - z * M.xy, g12 = x * M.zz - z * M.xz;
  const double g20 = y * M.xx - x * M.xy, g21 = y * M.xy - x * M.yy, g22 = y * M.xz - x * M.yz;
  // H_rr = G S (symmetric)
  DDLO_C(0) = z * g01 - y * g02, DDLO_C(1) = x * g02 - z * g00, DDLO_C(2) = y * g00 - x * g01;
  DDLO_C(3) = x * g12 - z * g10, DDLO_C(4) = y * g10 - x * g11, DDLO_C(5) = y * g20 - x * g21;
  // H_rt = -G
  DDLO_C(6) = -g00, DDLO_C(7) = -g01, DDLO_C(8) = -g02, DDLO_C(9) = -g10, DDLO_C(10) = -g11, DDLO_C(11) = -g12;
  DDLO_C(12) = -g20, DDLO_C(13) = -g21, DDLO_C(14) = -g22;
  // H_tt = M
  DDLO_C(15) = M.xx, DDLO_C(16) = M.xy, DDLO_C(17) = M.xz, DDLO_C(18) = M.yy, DDLO_C(19) = M.yz, DDLO_C(20) = M.zz;
  // b_r = G e, b_t = -M e
  DDLO_C(21) = g00 * ex + g01 * ey + g02 * ez;
  DDLO_C(22) = g10 * ex + g11 * ey + g12 * ez;
  DDLO_C(23) = g20 * ex + g21 * ey + g22 * ez;
  DDLO_C(24) = -mex, DDLO_C(25) = -mey, DDLO_C(26) = -mez;
#undef DDLO_C
}

// one source point of compute_error: stored correspondence and Mahalanobis matrix, new transform
__device__ __forceinline__ double err_point(const GicpArgs& a, const Iso3& T, int i) {
  const int j = __ldcg(a.corr + i);
  if (j < 0) return 0.0;
  const float4 pa = __ldg(a.src_pts + i);
  const float4 pb = __ldg(a.tgt.spts + __ldcg(a.nn_seed + i).x);
  const Sym3 M = load_sym3_cg(a.mahal + (size_t)i * kCovStride);
  const double x = xform_d(T.r + 0, T.t[0], (double)pa.x, (double)pa.y, (double)pa.z);
  const double y = xform_d(T.r + 3, T.t[1], (double)pa.x, (double)pa.y, (double)pa.z);
  const double z = xform_d(T.r + 6, T.t[2], (double)pa.x, (double)pa.y, (double)pa.z);
  const double ex = (double)pb.x - x, ey = (double)pb.y - y, ez = (double)pb.z - z;
  double mx, my, mz;
  return quad_form(M, ex, ey, ez, mx, my, mz);
}

// Work distribution.  The source points are dealt to the blocks sixteen consecutive points at a
// time (the queries of one warp: neighbours in the scan, so their searches walk the same nodes),
// round-robin, so that every block holds the same mix of cheap and expensive queries and the blocks
// reach the grid barrier together.  The assignment is static and all sums are taken per slot in a
// fixed order, so every bit of H, b and the error is reproducible run to run.
struct Deal {
  int nslots;  // slots of one block (multiple of 16)
  int nb, b;
  __device__ __forceinline__ int point(int slot) const { return (((slot >> 4) * nb + b) << 4) + (slot & 15); }
};
__device__ __forceinline__ Deal make_deal(int ns) {
  Deal d;
  d.nb = gridDim.x;
  d.b = blockIdx.x;
  d.nslots = 16 * (((ns + 15) / 16 + d.nb - 1) / d.nb);
  return d;
}

// Phase A of one round: update_correspondences' 1-NN search for the slots [base, base + nround) of
// this block.  Executed by the first kSearchWarps warps.  A pair of lanes serves one query at a
// time and fetches the next one from the round's queue (sm.next) as soon as it is done, so that a
// long search holds back neither the other pairs of its warp nor the block.  Which pair serves a
// slot never shows in the result: matches are parked per slot.
__device__ __forceinline__ void search_round(const GicpArgs& a, AlignSmem& sm, const Deal& dl, int base, int nround, bool have_prev) {
  const int lane = threadIdx.x & 31, h = lane & 1;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  Best1Pair best;
  unsigned node = 0;   // node to visit next
  unsigned start = 0;  // root of the subtree being searched
  int skip = -1, t = -1, sp = 0;
  int seed_start = 0, seed_count = 0;  // leaf to seed a fresh query from (pair-uniform), consumed right after the fetch
  bool run = false, sub_done = false, exhausted = false;
  unsigned long long stk[kPairStack];
#ifdef DDLO_VISIT_STATS
  int n_vis = 0, n_steps = 0;
#endif
  for (;;) {
    if (__any_sync(kFull, !run)) {
      if (sub_done) {
        // The subtree below `node` is searched.  Done if the ball of the best distance lies inside
        // the node's cube (always true at the root); else continue with the rest of the parent.
        sub_done = false;
        const int4 m = __ldg(a.tgt.meta + start);
        if (ball_in_cell(make_ball(a.tgt, qx, qy, qz, best.d), m)) {
          if (h == 0) {
            sm.nn_d[t] = best.d;
            sm.nn_idx[t] = best.idx;
            sm.nn_pos[t] = best.pos;
#ifdef DDLO_VISIT_STATS
            if (a.dbg_visits && sm.lm.n_lin < 4) a.dbg_visits[(size_t)sm.lm.n_lin * a.ns + dl.point(base + t)] = make_int4(n_vis, 0, n_steps, 0);
#endif
          }
        } else {
          skip = (int)start;
          start = node = (unsigned)m.x;
          sp = 0;
          run = true;
        }
      }
      const bool need = !run && !exhausted;
      const unsigned nm = __ballot_sync(kFull, need && h == 0);
      if (nm) {
        const int leader = __ffs(nm) - 1;
        int first = 0;
        if (lane == leader) first = atomicAdd(&sm.next, __popc(nm));
        first = __shfl_sync(kFull, first, leader);
        if (need) {
          t = first + __popc(nm & ((1u << (lane & ~1)) - 1u));
          if (t >= nround) {
            exhausted = true;
          } else {
            const int i = dl.point(base + t);
            best = Best1Pair();
            if (i < a.ns && a.tgt.n > 0) {
              const float4 pa = __ldg(a.src_pts + i);
              qx = xform_f(sm.lm.Rf + 0, sm.lm.tf[0], pa.x, pa.y, pa.z);
              qy = xform_f(sm.lm.Rf + 3, sm.lm.tf[1], pa.x, pa.y, pa.z);
              qz = xform_f(sm.lm.Rf + 6, sm.lm.tf[2], pa.x, pa.y, pa.z);
              start = 0;
              skip = -1;
              sp = 0;
              const int2 sd = have_prev ? __ldcg(a.nn_seed + i) : make_int2(-1, -1);
              if (sd.x >= 0) {
                // The previous iteration's match is a real point of the target: its distance bounds the
                // answer, and the search starts at the node above its leaf and climbs only while the
                // ball of the best distance found so far sticks out of the node's cube.
                const float4 tp = __ldg(a.tgt.spts + sd.x);
                best.seed(sqdist3_rn(qx, qy, qz, tp.x, tp.y, tp.z), __float_as_int(tp.w), sd.x);
                start = (unsigned)sd.y;
              } else {
                // No previous match: walk down the cells that contain the query itself, one child
                // reference per level (no boxes), to the leaf it falls into.  That leaf's points seed the
                // search, which then starts at the node above it exactly like a seeded one.  An empty
                // slot on the way just means: start at that node without a seed.
                const float4 lat = __ldg(reinterpret_cast<const float4*>(a.tgt.lattice));
                const unsigned cx = (unsigned)fminf(fmaxf((qx - lat.x) * lat.w, 0.0f), 1023.0f);
                const unsigned cy = (unsigned)fminf(fmaxf((qy - lat.y) * lat.w, 0.0f), 1023.0f);
                const unsigned cz = (unsigned)fminf(fmaxf((qz - lat.z) * lat.w, 0.0f), 1023.0f);
                unsigned nd = 0;
                for (int sh = kMortonLevels - 1; sh >= 0; --sh) {
                  const int slot = (int)(((cx >> sh) & 1u) | (((cy >> sh) & 1u) << 1) | (((cz >> sh) & 1u) << 2));
                  const int2 ref = __ldg(reinterpret_cast<const int2*>(a.tgt.nodes + (size_t)nd * kNodeF4 + 12) + slot);
                  if (ref.y < 0) {
                    nd = (unsigned)ref.x;
                    continue;
                  }
                  if (ref.y > 0) {
                    seed_start = ref.x;
                    seed_count = ref.y;
                  }
                  break;
                }
                start = nd;
              }
              node = start;
              run = true;
#ifdef DDLO_VISIT_STATS
              n_vis = 0;
              n_steps = 0;
#endif
            } else if (h == 0) {  // padding slot of the last group, or an empty target: no match
              sm.nn_d[t] = FLT_MAX;
              sm.nn_idx[t] = kIdxSentinel;
              sm.nn_pos[t] = -1;
            }
          }
        }
      }
    }
    if (__any_sync(kFull, seed_count > 0)) {  // new queries without a previous match: seed from their own leaf
      best.scan(seed_count > 0, a.tgt.spts, seed_start, seed_count, qx, qy, qz, h);
      seed_count = 0;
    }
    if (!__any_sync(kFull, run)) {
      if (__all_sync(kFull, exhausted)) break;
      continue;  // padding slots only: fetch again
    }
#ifdef DDLO_VISIT_STATS
    n_vis += run ? 1 : 0;
    n_steps += 1;
#endif
    const bool was = run;
    nn1_visit_pair(a.tgt, run, qx, qy, qz, best, node, skip, stk, sp, h);
    sub_done = was && !run;
  }
}

// linearize over the block's points; leaves the block's 28 sums in dst[c * stride + blockIdx.x].
// have_prev: nn_seed holds the matches of the previous linearize of the same source cloud.
__device__ __forceinline__ void linearize_block(const GicpArgs& a, AlignSmem& sm, bool have_prev, double* dst, int stride,
                                                unsigned long long* bt = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kNumSums) sm.acc[threadIdx.x] = 0.0;
  if (threadIdx.x == 0) sm.t_search = 0ull;
  const Deal dl = make_deal(a.ns);
  for (int base = 0; base < dl.nslots; base += kAlignPairs) {
    const int nround = min(kAlignPairs, dl.nslots - base);  // multiple of 16
    if (threadIdx.x == 0) sm.next = 0;
    __syncthreads();  // queue reset; parked matches and contrib of the previous round consumed; acc initialised
    // ---- phase A
    if (warp < kSearchWarps) {
      search_round(a, sm, dl, base, nround, have_prev);
      if (bt && lane == 0) atomicMax(&sm.t_search, globaltimer_ns());  // when the block's last warp left the search
    }
    __syncthreads();
    // ---- phase B: one thread per point of the round
    if (threadIdx.x < nround) {
      const int i = dl.point(base + threadIdx.x);
      const unsigned long long t0 = bt ? globaltimer_ns() : 0ull;
      lin_point(a, sm.lm, i < a.ns, i, sm.nn_d[threadIdx.x], sm.nn_idx[threadIdx.x], sm.nn_pos[threadIdx.x], sm.contrib + threadIdx.x);
      if (bt && lane == 0) {
        const unsigned long long dt = globaltimer_ns() - t0;
        atomicMax(bt + 4, dt);
        atomicMax(bt + 5, 0xffffffffull - dt);
      }
    }
    __syncthreads();
    // ---- block sum of this round, fixed order: warp c adds component c over the round's slots
    if (warp < kNumSums) {
      double v = 0.0;
      for (int p = lane; p < nround; p += 32) v += sm.contrib[warp * kAlignPairs + p];
      v = warp_sum(v);
      if (lane == 0) sm.acc[warp] += v;
    }
  }
  __syncthreads();
  if (bt && threadIdx.x == 0) bt[6] = globaltimer_ns();
  if (threadIdx.x < kNumSums) __stcg(dst + (size_t)threadIdx.x * stride + blockIdx.x, sm.acc[threadIdx.x]);
  __syncthreads();
  if (bt && threadIdx.x == 0) bt[7] = globaltimer_ns();
}

// compute_error over the block's points; the block's sum goes to dst[blockIdx.x]
__device__ __forceinline__ void error_block(const GicpArgs& a, const Iso3& T, AlignSmem& sm, double* dst) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Deal dl = make_deal(a.ns);
  double e = 0.0;
  for (int slot = threadIdx.x; slot < dl.nslots; slot += kAlignThreads) {
    const int i = dl.point(slot);
    if (i < a.ns) e += err_point(a, T, i);
  }
  e = warp_sum(e);
  if (lane == 0) sm.red[warp] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < kAlignWarps; ++w) v += sm.red[w];
    __stcg(dst + blockIdx.x, v);
  }
  __syncthreads();
}

// grid: every block sums all per-block partials in the same fixed order (L2 reads, L1 bypassed)
template <int NCOMP>
__device__ __forceinline__ void grid_sum(const double* src, int stride, int nblk, double* s_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c = warp; c < NCOMP; c += nwarp) {
    double v = 0.0;
    for (int b = lane; b < nblk; b += 32) v += __ldcg(src + (size_t)c * stride + b);
    v = warp_sum(v);
    if (lane == 0) s_tot[c] = v;
  }
  __syncthreads();
}

__device__ __forceinline__ void unpack_sums(const double* t, double* H /*row-major 6x6*/, double* b, double& err) {
  H[0] = t[0], H[1] = t[1], H[2] = t[2], H[7] = t[3], H[8] = t[4], H[14] = t[5];
  H[6] = t[1], H[12] = t[2], H[13] = t[4];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      H[6 * r + 3 + c] = t[6 + 3 * r + c];
      H[6 * (3 + c) + r] = t[6 + 3 * r + c];
    }
  H[21] = t[15], H[22] = t[16], H[23] = t[17], H[28] = t[18], H[29] = t[19], H[35] = t[20];
  H[27] = t[16], H[33] = t[17], H[34] = t[19];
  for (int r = 0; r < 6; ++r) b[r] = t[21 + r];
  err = t[27];
}

// lsq_registration_impl.hpp:129-139
__device__ __forceinline__ bool is_converged(const Iso3& delta, double rot_eps, double trans_eps) {
  double rmax = 0.0, tmax = 0.0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) rmax = fmax(rmax, 1.0 / rot_eps * fabs(delta.r[3 * i + j] - (i == j ? 1.0 : 0.0)));
    tmax = fmax(tmax, 1.0 / trans_eps * fabs(delta.t[i]));
  }
  return fmax(rmax, tmax) < 1;
}

__device__ __forceinline__ void iso_from_colmajor(const float* m, Iso3& T) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T.r[3 * i + j] = (double)m[4 * j + i];
    T.t[i] = (double)m[12 + i];
  }
}
__device__ __forceinline__ void iso_from_colmajor(const double* m, Iso3& T) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T.r[3 * i + j] = m[4 * j + i];
    T.t[i] = m[12 + i];
  }
}

// solve (H + lambda I) d = -b, delta = [exp(d_0..2) | d_3..5]
__device__ __noinline__ void lm_solve(LmShared& s, double lambda) {
  double A[36], nb[6];
  for (int i = 0; i < 36; ++i) A[i] = s.H[i];
  for (int i = 0; i < 6; ++i) {
    A[7 * i] += lambda;
    nb[i] = -s.b[i];
  }
  ldlt6_solve_fast(A, nb, s.d);
  so3_exp_matrix(s.d, s.delta.r);
  for (int i = 0; i < 3; ++i) s.delta.t[i] = s.d[3 + i];
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

  if (threadIdx.x == 0) {
    iso_from_colmajor(a.guess, s.x0);
    s.lambda = -1.0;  // lm_lambda_ = -1 (:100)
    s.converged = 0;
    s.lm_failed = 0;
    s.n_lin = s.n_err = 0;
    s.nr_iter = 0;
    s.final_error = 0.0;
    for (int i = 0; i < 36; ++i) s.final_H[i] = (i % 7 == 0) ? 1.0 : 0.0;  // final_hessian_.setIdentity()
  }
  __syncthreads();

  for (int it = 0; it < a.max_iterations; ++it) {
    if (s.converged) break;
    __syncthreads();
    if (threadIdx.x == 0) {
      s.nr_iter = it;
      iso_to_float(s.x0, s.Rf, s.tf);
    }
    __syncthreads();

    // ---- linearize(x0) -----------------------------------------------------------------------
    double* part = a.partials + (size_t)(seq & 1) * kNumSums * a.partial_stride;
    unsigned long long* bt = (a.blk_times && it < 8) ? a.blk_times + ((size_t)it * nblk + blockIdx.x) * 8 : nullptr;
    if (bt && threadIdx.x == 0) bt[0] = globaltimer_ns();
    linearize_block(a, sm, it > 0, part, a.partial_stride, bt);
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
    ++seq;
    DDLO_STAMP(kTagLinSummed);

    if (threadIdx.x == 0) {
      unpack_sums(sm.tot, s.H, s.b, s.y0);
      s.n_lin += 1;
      s.step_ok = 0;
      if (a.optimizer == DDLO_OPT_GAUSS_NEWTON) {
        // step_gn (:156-173)
        lm_solve(s, 0.0);
        s.x0 = iso_mul(s.delta, s.x0);
        for (int i = 0; i < 36; ++i) s.final_H[i] = s.H[i];
        s.final_error = s.y0;
        s.step_ok = 1;
      } else {
        if (s.lambda < 0.0) {
          double mx = 0.0;
          for (int i = 0; i < 6; ++i) mx = fmax(mx, fabs(s.H[7 * i]));
          s.lambda = a.lm_init_lambda_factor * mx;
        }
        s.nu = 2.0;
      }
    }
    __syncthreads();

    if (a.optimizer != DDLO_OPT_GAUSS_NEWTON) {
      // ---- step_lm trials (:188-229) ---------------------------------------------------------
      for (int trial = 0; trial < a.lm_max_iterations; ++trial) {
        if (threadIdx.x == 0) {
          lm_solve(s, s.lambda);
          s.xi = iso_mul(s.delta, s.x0);
        }
        __syncthreads();
        DDLO_STAMP(kTagSolved);
        double* epart = a.partials + (size_t)(seq & 1) * kNumSums * a.partial_stride;
        error_block(a, s.xi, sm, epart);
        DDLO_STAMP(kTagErrDone);
        grid.sync();
        DDLO_STAMP(kTagErrSynced);
        grid_sum<1>(epart, a.partial_stride, nblk, sm.tot);
        ++seq;
        if (threadIdx.x == 0) {
          s.n_err += 1;
          s.yi = sm.tot[0];
          double den = 0.0;
          for (int r = 0; r < 6; ++r) den += s.d[r] * (s.lambda * s.d[r] - s.b[r]);
          const double rho = (s.y0 - s.yi) / den;
          if (rho < 0) {
            if (is_converged(s.delta, a.rot_eps, a.trans_eps)) {
              s.step_ok = 1;  // returns true with x0 unchanged (:215-218)
              s.action = 1;
            } else {
              s.lambda = s.nu * s.lambda;
              s.nu = 2 * s.nu;
              s.action = 0;
            }
          } else {  // also taken when rho is NaN, like the reference's `if (rho < 0)`
            s.x0 = s.xi;
            const double q = 2 * rho - 1;
            s.lambda = s.lambda * fmax(1.0 / 3.0, 1 - q * q * q);
            for (int i = 0; i < 36; ++i) s.final_H[i] = s.H[i];
            s.final_error = s.yi;
            s.step_ok = 1;
            s.action = 1;
          }
        }
        __syncthreads();
        DDLO_STAMP(kTagDecided);
        if (s.action) break;
      }
    }
    __syncthreads();
    if (!s.step_ok) {  // "lm not converged!!" (:115-119)
      if (threadIdx.x == 0) s.lm_failed = 1;
      __syncthreads();
      break;
    }
    if (threadIdx.x == 0) s.converged = is_converged(s.delta, a.rot_eps, a.trans_eps) ? 1 : 0;
    __syncthreads();
  }
  __syncthreads();

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    AlignOut* o = a.out;
    for (int i = 0; i < 16; ++i) o->final_transformation[i] = 0.0f;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) o->final_transformation[4 * j + i] = (float)s.x0.r[3 * i + j];
      o->final_transformation[12 + i] = (float)s.x0.t[i];
    }
    o->final_transformation[15] = 1.0f;
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) o->final_hessian[6 * j + i] = s.final_H[6 * i + j];
    const bool bad = reinterpret_cast<const int*>(a.tgt.lattice)[4] != 0 || (a.src_lattice && reinterpret_cast<const int*>(a.src_lattice)[4] != 0);
    o->flags = (s.converged ? DDLO_FLAG_CONVERGED : 0) | (s.lm_failed ? DDLO_FLAG_LM_FAILED : 0) | (bad ? DDLO_FLAG_NONFINITE : 0);
    o->nr_iterations = s.nr_iter;
    o->n_linearize = s.n_lin;
    o->n_compute_error = s.n_err;
    o->final_error = s.final_error;
    o->lm_lambda = s.lambda;
    DDLO_STAMP(kTagEnd);
    if (a.stamps) a.stamps[0] = (unsigned long long)n_stamps;
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
