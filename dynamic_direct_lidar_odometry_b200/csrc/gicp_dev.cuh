// Device code shared by the single-registration kernel (gicp.cu: k_align, one cooperative launch per align) and the
// batched kernels (batch_align.cu: many registrations per launch):
//
//   search_slots     NanoGICP::update_correspondences' 1-NN search for a range of a chunk's slots
//                                                                      (nano_gicp_impl.hpp:235-275)
//   lin_group        the body of linearize for 32 consecutive slots     (nano_gicp_impl.hpp:292-328)
//   err_point        the body of compute_error                          (nano_gicp_impl.hpp:349-368)
//   lm_start / lm_on_linearized / lm_on_error
//                    LsqRegistration::computeTransformation, step_lm, step_gn, is_converged as a state machine
//                                                                      (lsq_registration_impl.hpp:96-232)
//
// Work decomposition and summation order (the same in both kernels, which makes their results bit-identical).
// The source points are dealt 16 at a time, round-robin, to `nb` CHUNKS (Deal); a chunk's slots are taken in GROUPS of
// 32 consecutive slots.  A group's 28 sums (J^T M J, J^T M e, e^T M e) are butterfly sums over the 32 lanes of the
// warp that serves it; a chunk's sums are the group sums added in group order; the registration's sums are the chunk
// sums added lane-strided and then by butterfly (grid_sum).  None of this depends on which warp, block or SM does the
// work, nor on the block size.
#pragma once

#include "gicp.cuh"
#include "knn.cuh"
#include "knn_pair.cuh"

namespace ddlo {

// sum layout: [0..5] H_rr upper, [6..14] H_rt row-major, [15..20] H_tt upper, [21..23] b_r,
// [24..26] b_t, [27] sum of e^T M e
struct LmShared {
  Iso3 x0, xi, delta;
  float Rf[9], tf[3];  // float cast of the transform used for the 1-NN queries (:240)
  double H[36], b[6], d[6];
  double y0, yi, lambda, nu, final_error;
  double final_H[36];
  int it, trial;
  int converged, step_ok, lm_failed, n_lin, n_err, nr_iter;
};

enum { kNextLinearize = 0, kNextError = 1, kNextDone = 2 };

constexpr int kGroup = 32;  // slots per group = lanes of the warp that sums them

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

// Work distribution.  The source points are dealt to `nb` chunks sixteen consecutive points at a time (the queries of
// one warp: neighbours in the scan, so their searches walk the same nodes), round-robin, so that every chunk holds the
// same mix of cheap and expensive queries.  The assignment is static and all sums are taken per slot in a fixed order,
// so every bit of H, b and the error is reproducible run to run.
struct Deal {
  int nslots;  // slots of one chunk (multiple of 16)
  int nb, b;
  __device__ __forceinline__ int point(int slot) const { return (((slot >> 4) * nb + b) << 4) + (slot & 15); }
};
__device__ __forceinline__ Deal make_deal(int ns, int nb, int b) {
  Deal d;
  d.nb = nb;
  d.b = b;
  d.nslots = 16 * (((ns + 15) / 16 + d.nb - 1) / d.nb);
  return d;
}

// What update_correspondences leaves per source point (:255-274): sq_distances_, correspondences_ and, for the next
// search, where the match sits in the Morton order and the node above its leaf.  Returns the correspondence.
__device__ __forceinline__ int store_match(const GicpArgs& a, int i, float nn_d, int nn_idx, int nn_pos) {
  DDLO_CHECK_INDEX(i, a.ns, "store_match: source index");
  a.sqd[i] = nn_d;
  const bool found = nn_idx != kIdxSentinel;
  if (found) {
    DDLO_CHECK_INDEX(nn_idx, a.tgt.n, "store_match: matched target index");
    DDLO_CHECK_INDEX(nn_pos, a.tgt.n, "store_match: matched Morton position");
  }
  const int j = (found && (double)nn_d < a.thr2) ? nn_idx : -1;
  a.corr[i] = j;
  // seed of the next search (kept even beyond the distance threshold)
  a.nn_seed[i] = found ? make_int2(nn_pos, __ldg(a.tgt.node_of_point + nn_idx)) : make_int2(-1, -1);
  return j;
}

// linearize for the 32 slots of one group, one slot per lane (MUST be executed by all 32 lanes; i < 0 or j < 0: the
// lane has no point / no correspondence and contributes zeros).  Writes the Mahalanobis matrix of the lane's point and
// returns, in lane c < 28, the group's sum of component c.
// Kept out of line so that its fp64 register appetite does not leak into the search loop.
static __device__ __noinline__ double lin_group(const GicpArgs& a, const Iso3& x0, int i, int j, int nn_pos) {
  const int lane = threadIdx.x & 31;
  const bool on = i >= 0 && j >= 0;
  double x = 0.0, y = 0.0, z = 0.0, ex = 0.0, ey = 0.0, ez = 0.0;
  Sym3 M{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (on) {
    DDLO_CHECK_INDEX(i, a.ns, "lin_group: source index");
    DDLO_CHECK_INDEX(nn_pos, a.tgt.n, "lin_group: target position");
    const float4 pa = __ldg(a.src_pts + i);
    const float4 pb = __ldg(a.tgt.spts + nn_pos);
    const Sym3 CA = load_sym3(a.src_cov + (size_t)i * kCovStride);
    const Sym3 CB = load_sym3(a.tgt_cov + (size_t)nn_pos * kCovStride);
    Sym3 RCR = sym3_rotate(x0.r, CA);
    RCR.xx += CB.xx, RCR.xy += CB.xy, RCR.xz += CB.xz, RCR.yy += CB.yy, RCR.yz += CB.yz, RCR.zz += CB.zz;
    M = sym3_inverse(RCR);
    store_sym3(a.mahal + (size_t)i * kCovStride, M);
    x = xform_d(x0.r + 0, x0.t[0], (double)pa.x, (double)pa.y, (double)pa.z);
    y = xform_d(x0.r + 3, x0.t[1], (double)pa.x, (double)pa.y, (double)pa.z);
    z = xform_d(x0.r + 6, x0.t[2], (double)pa.x, (double)pa.y, (double)pa.z);
    ex = (double)pb.x - x, ey = (double)pb.y - y, ez = (double)pb.z - z;
  }
  double mine = 0.0;
#define DDLO_C(k, expr)                  \
  do {                                   \
    const double s__ = warp_sum(expr);   \
    if (lane == (k)) mine = s__;         \
  } while (0)
  double mex, mey, mez;
  const double q = quad_form(M, ex, ey, ez, mex, mey, mez);
  DDLO_C(27, q);
  // G = S^T M with S = skew(T p_A);  J = [S | -I]
  const double g00 = z * M.xy - y * M.xz, g01 = z * M.yy - y * M.yz, g02 = z * M.yz - y * M.zz;
  const double g10 = x * M.xz - z * M.xx, g11 = x * M.yz - z * M.xy, g12 = x * M.zz - z * M.xz;
  const double g20 = y * M.xx - x * M.xy, g21 = y * M.xy - x * M.yy, g22 = y * M.xz - x * M.yz;
  // H_rr = G S (symmetric)
  DDLO_C(0, z * g01 - y * g02);
  DDLO_C(1, x * g02 - z * g00);
  DDLO_C(2, y * g00 - x * g01);
  DDLO_C(3, x * g12 - z * g10);
  DDLO_C(4, y * g10 - x * g11);
  DDLO_C(5, y * g20 - x * g21);
  // H_rt = -G
  DDLO_C(6, -g00);
  DDLO_C(7, -g01);
  DDLO_C(8, -g02);
  DDLO_C(9, -g10);
  DDLO_C(10, -g11);
  DDLO_C(11, -g12);
  DDLO_C(12, -g20);
  DDLO_C(13, -g21);
  DDLO_C(14, -g22);
  // H_tt = M
  DDLO_C(15, M.xx);
  DDLO_C(16, M.xy);
  DDLO_C(17, M.xz);
  DDLO_C(18, M.yy);
  DDLO_C(19, M.yz);
  DDLO_C(20, M.zz);
  // b_r = G e, b_t = -M e
  DDLO_C(21, g00 * ex + g01 * ey + g02 * ez);
  DDLO_C(22, g10 * ex + g11 * ey + g12 * ez);
  DDLO_C(23, g20 * ex + g21 * ey + g22 * ez);
  DDLO_C(24, -mex);
  DDLO_C(25, -mey);
  DDLO_C(26, -mez);
#undef DDLO_C
  return mine;
}

// compute_error for the 32 slots of one group, one slot per lane (all 32 lanes; i < 0: no point): stored
// correspondence and Mahalanobis matrix, new transform.  Returns the group's sum in every lane.
// Out of line for the same reason as lm_on_*: one compiled body for every kernel that uses it.
static __device__ __noinline__ double err_group(const GicpArgs& a, const Iso3& T, int i);

// one source point of compute_error
__device__ __forceinline__ double err_point(const GicpArgs& a, const Iso3& T, int i) {
  if (i < 0) return 0.0;
  DDLO_CHECK_INDEX(i, a.ns, "err_point: source index");
  // every load that depends on i alone is issued before the first use (the pass is a stream of 84 bytes per point:
  // what it needs is loads in flight); only the matched target point waits for the stored position
  const int j = __ldcg(a.corr + i);
  const int2 seed = __ldcg(a.nn_seed + i);
  const float4 pa = __ldg(a.src_pts + i);
  const Sym3 M = load_sym3_cg(a.mahal + (size_t)i * kCovStride);  // (stale for a point without correspondence: not used then)
  if (j < 0) return 0.0;
  DDLO_CHECK_INDEX(seed.x, a.tgt.n, "err_point: stored target position");
  const float4 pb = __ldg(a.tgt.spts + seed.x);
  const double x = xform_d(T.r + 0, T.t[0], (double)pa.x, (double)pa.y, (double)pa.z);
  const double y = xform_d(T.r + 3, T.t[1], (double)pa.x, (double)pa.y, (double)pa.z);
  const double z = xform_d(T.r + 6, T.t[2], (double)pa.x, (double)pa.y, (double)pa.z);
  const double ex = (double)pb.x - x, ey = (double)pb.y - y, ez = (double)pb.z - z;
  double mx, my, mz;
  return quad_form(M, ex, ey, ez, mx, my, mz);
}

static __device__ __noinline__ double err_group(const GicpArgs& a, const Iso3& T, int i) { return warp_sum(err_point(a, T, i)); }

// Where a search round parks its matches (shared memory of the block that runs it), its queue head, and the float
// transform of the queries.
struct SearchPark {
  float* nn_d;
  int* nn_idx;
  int* nn_pos;
  int* next;  // queue head
  const float* Rf;
  const float* tf;
  int n_lin;  // DDLO_VISIT_STATS builds: pass index
};

// update_correspondences' 1-NN search for the slots [base, base + nround) of a chunk.  Executed by whole warps (any
// number of them).  A pair of lanes serves one query at a time and fetches the next one from the round's queue
// (*pk.next) as soon as it is done, so that a long search holds back neither the other pairs of its warp nor the
// block.  Which pair serves a slot never shows in the result: matches are parked per slot (pk.nn_*[slot - base]).
__device__ __forceinline__ void search_slots(const GicpArgs& a, const SearchPark& pk, const Deal& dl, int base, int nround, bool have_prev) {
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
            pk.nn_d[t] = best.d;
            pk.nn_idx[t] = best.idx;
            pk.nn_pos[t] = best.pos;
#ifdef DDLO_VISIT_STATS
            if (a.dbg_visits && pk.n_lin < 4) a.dbg_visits[(size_t)pk.n_lin * a.ns + dl.point(base + t)] = make_int4(n_vis, 0, n_steps, 0);
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
        if (lane == leader) first = atomicAdd(pk.next, __popc(nm));
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
              qx = xform_f(pk.Rf + 0, pk.tf[0], pa.x, pa.y, pa.z);
              qy = xform_f(pk.Rf + 3, pk.tf[1], pa.x, pa.y, pa.z);
              qz = xform_f(pk.Rf + 6, pk.tf[2], pa.x, pa.y, pa.z);
              start = 0;
              skip = -1;
              sp = 0;
              const int2 sd = have_prev ? __ldcg(a.nn_seed + i) : make_int2(-1, -1);
              if (sd.x >= 0) {
                // The previous iteration's match is a real point of the target: its distance bounds the
                // answer, and the search starts at the node above its leaf and climbs only while the
                // ball of the best distance found so far sticks out of the node's cube.
                DDLO_CHECK_INDEX(sd.x, a.tgt.n, "search: seed position");
                DDLO_CHECK_INDEX(sd.y, reinterpret_cast<const int*>(a.tgt.lattice)[5], "search: seed node");
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
              pk.nn_d[t] = FLT_MAX;
              pk.nn_idx[t] = kIdxSentinel;
              pk.nn_pos[t] = -1;
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

// every block sums all per-chunk partials in the same fixed order (L2 reads, L1 bypassed)
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
static __device__ __noinline__ void lm_solve(LmShared& s, double lambda) {
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

// ---- the LM / GN driver as a state machine (one thread) ------------------------------------------------------------
// computeTransformation (:96-127): x0 = guess, lm_lambda_ = -1, then up to max_iterations_ outer iterations, each
// a linearize followed (LM) by up to lm_max_iterations_ trial errors.  Each function returns what has to be
// evaluated next: linearize at x0 (queries use Rf/tf), compute_error at xi, or nothing (done).
// (Out of line on purpose: k_align and the batched kernels must run the very same instruction sequence - an inlined
// copy may contract its multiply-adds differently in each kernel, and the results are promised to be bit-identical.)
static __device__ __noinline__ int lm_start(LmShared& s, const GicpArgs& a) {
  iso_from_colmajor(a.guess, s.x0);
  s.lambda = -1.0;  // lm_lambda_ = -1 (:100)
  s.converged = 0;
  s.lm_failed = 0;
  s.n_lin = s.n_err = 0;
  s.nr_iter = 0;
  s.it = 0;
  s.trial = 0;
  s.final_error = 0.0;
  for (int i = 0; i < 36; ++i) s.final_H[i] = (i % 7 == 0) ? 1.0 : 0.0;  // final_hessian_.setIdentity()
  if (a.max_iterations <= 0) return kNextDone;
  iso_to_float(s.x0, s.Rf, s.tf);
  return kNextLinearize;
}

// the end of one outer iteration (:113-122)
static __device__ __noinline__ int lm_end_iteration(LmShared& s, const GicpArgs& a) {
  if (!s.step_ok) {  // "lm not converged!!" (:115-119)
    s.lm_failed = 1;
    return kNextDone;
  }
  s.converged = is_converged(s.delta, a.rot_eps, a.trans_eps) ? 1 : 0;
  s.it += 1;
  if (s.converged || s.it >= a.max_iterations) return kNextDone;
  s.nr_iter = s.it;
  iso_to_float(s.x0, s.Rf, s.tf);
  return kNextLinearize;
}

// after linearize(x0): tot = the 28 sums
static __device__ __noinline__ int lm_on_linearized(LmShared& s, const GicpArgs& a, const double* tot) {
  unpack_sums(tot, s.H, s.b, s.y0);
  s.n_lin += 1;
  s.step_ok = 0;
  if (a.optimizer == DDLO_OPT_GAUSS_NEWTON) {
    // step_gn (:156-173)
    lm_solve(s, 0.0);
    s.x0 = iso_mul(s.delta, s.x0);
    for (int i = 0; i < 36; ++i) s.final_H[i] = s.H[i];
    s.final_error = s.y0;
    s.step_ok = 1;
    return lm_end_iteration(s, a);
  }
  // step_lm (:176-232)
  if (s.lambda < 0.0) {
    double mx = 0.0;
    for (int i = 0; i < 6; ++i) mx = fmax(mx, fabs(s.H[7 * i]));
    s.lambda = a.lm_init_lambda_factor * mx;
  }
  s.nu = 2.0;
  s.trial = 0;
  if (a.lm_max_iterations <= 0) return lm_end_iteration(s, a);
  lm_solve(s, s.lambda);
  s.xi = iso_mul(s.delta, s.x0);
  return kNextError;
}

// after compute_error(xi) = yi (:199-229)
static __device__ __noinline__ int lm_on_error(LmShared& s, const GicpArgs& a, double yi) {
  s.n_err += 1;
  s.yi = yi;
  double den = 0.0;
  for (int r = 0; r < 6; ++r) den += s.d[r] * (s.lambda * s.d[r] - s.b[r]);
  const double rho = (s.y0 - s.yi) / den;
  if (rho < 0) {
    if (is_converged(s.delta, a.rot_eps, a.trans_eps)) {
      s.step_ok = 1;  // returns true with x0 unchanged (:215-218)
      return lm_end_iteration(s, a);
    }
    s.lambda = s.nu * s.lambda;
    s.nu = 2 * s.nu;
    s.trial += 1;
    if (s.trial >= a.lm_max_iterations) return lm_end_iteration(s, a);
    lm_solve(s, s.lambda);
    s.xi = iso_mul(s.delta, s.x0);
    return kNextError;
  }
  // also taken when rho is NaN, like the reference's `if (rho < 0)`
  s.x0 = s.xi;
  const double q = 2 * rho - 1;
  s.lambda = s.lambda * fmax(1.0 / 3.0, 1 - q * q * q);
  for (int i = 0; i < 36; ++i) s.final_H[i] = s.H[i];
  s.final_error = s.yi;
  s.step_ok = 1;
  return lm_end_iteration(s, a);
}

// what align() leaves for the host
static __device__ __noinline__ void write_align_out(const GicpArgs& a, const LmShared& s, AlignOut* o) {
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
}

}  // namespace ddlo
