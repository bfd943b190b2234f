// Exact nearest-neighbour search in the Morton-prefix octree, one query per PAIR of lanes.
//
// Used by update_correspondences (nano_gicp_impl.hpp:235-275 -> nanoflann_impl.hpp:1365-1384,
// 1495-1566 with k = 1, result set Best1Pair).  (A k = 20 result set over two lanes was tried for
// the covariances' self k-NN and lost to the 8-lane one of knn.cuh: with 16 queries behind a warp
// instruction some query accepts a candidate at nearly every step, so the insertion network runs all
// the time at low utilisation.)  knn.cuh serves a query with 8 lanes, one child box per lane; here a query
// owns 2 lanes and every lane bounds 4 children with 128-bit loads.  A warp therefore carries 16
// queries instead of 4, so that all the source points of one SM (about 443 on the 64x1024 scan) are
// in flight at once, and the bookkeeping instructions (votes, shuffles, stack handling), which cost
// the same whatever the number of queries behind a warp instruction, are shared by four times as
// many queries.  Pending internal children go on a small stack in the lane's LOCAL memory (each
// lane keeps the children it bounded itself; a pop takes the nearer of the two tops), which frees
// the shared memory the 8-lane version needed for its stacks.
//
// Exactness is argued exactly as in knn.cuh: distances and box bounds use the reference's float
// expression tree with round-to-nearest intrinsics, a subtree is skipped only if its bound exceeds
// the best distance found so far, and candidates are ranked by (d2, original index).
#pragma once

#include "knn.cuh"

namespace ddlo {

constexpr int kPairStack = 40;  // <= 4 pushes per lane and level, internal children live on levels 1..9

// (Measured in round 2: keeping the first 12 - 16 stack entries of every lane in shared memory instead of local memory
// leaves the DRAM traffic of k_align unchanged - its extra write traffic is the write-back of the preceding kernels'
// dirty lines, not the stacks - and costs 13 us of 265 for the address arithmetic on the search's critical chain.)

// best candidate of a pair of lanes (uniform across the pair)
struct Best1Pair {
  float d = FLT_MAX;
  int idx = kIdxSentinel;  // original index
  int pos = -1;            // position in spts (Morton order)
  __device__ __forceinline__ float worst() const { return d; }
  static __device__ __forceinline__ unsigned long long pack(float dist, int oi) {
    return ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)oi;
  }
  __device__ __forceinline__ void seed(float dist, int oi, int p) {
    if (dist < FLT_MAX && lex_less(dist, oi, d, idx)) {
      d = dist;
      idx = oi;
      pos = p;
    }
  }
  // all 32 lanes; `doit`, start and count uniform per pair; h = lane & 1
  __device__ __forceinline__ void scan(bool doit, const float4* __restrict__ spts, int start, int count, float qx, float qy, float qz,
                                       int h) {
    unsigned long long key = pack(d, idx);
    int lp = pos;
    if (doit) {
      DDLO_CHECK_INDEX(start + count - 1, 0x7fffffff, "Best1Pair::scan: leaf range");
#pragma unroll 4
      for (int j = h; j < count; j += 2) {
        const float4 v = __ldg(spts + start + j);
        const float dist = sqdist3_rn(qx, qy, qz, v.x, v.y, v.z);
        const unsigned long long k2 = pack(dist, __float_as_int(v.w));
        if (dist < FLT_MAX && k2 < key) {
          key = k2;
          lp = start + j;
        }
      }
    }
    const unsigned long long ok = __shfl_xor_sync(kFull, key, 1);
    const int olp = __shfl_xor_sync(kFull, lp, 1);
    if (ok < key) {
      key = ok;
      lp = olp;
    }
    pos = lp;
    d = __uint_as_float((unsigned)(key >> 32));
    idx = (int)(unsigned)(key & 0xffffffffull);
  }
};

__device__ __forceinline__ float sel4(const float (&a)[4], int c) { return c == 0 ? a[0] : (c == 1 ? a[1] : (c == 2 ? a[2] : a[3])); }
__device__ __forceinline__ int sel4(const int (&a)[4], int c) { return c == 0 ? a[0] : (c == 1 ? a[1] : (c == 2 ? a[2] : a[3])); }

// One node visit of the depth-first search of a pair of lanes.  MUST be executed by all 32 lanes;
// `run`, the query, rs, node and skip are uniform per pair; stk/sp are the lane's own pending
// children.  `skip` (or -1) is a child NODE of the subtree's start node that has been searched
// already.  On return `node` is the next node to visit, or run == false when the subtree is done.
template <class RS>
__device__ __forceinline__ void nn1_visit_pair(const IndexView& ix, bool& run, float qx, float qy, float qz, RS& rs, unsigned& node,
                                               int skip, unsigned long long* __restrict__ stk, int& sp, int h) {
  const float inf = __int_as_float(0x7f800000);
  float b[4] = {inf, inf, inf, inf};
  int rx[4] = {0, 0, 0, 0}, ry[4] = {0, 0, 0, 0};
  if (run) {
    DDLO_CHECK_INDEX(node, reinterpret_cast<const int*>(ix.lattice)[5], "nn1_visit_pair: node");
    const float4* g = ix.nodes + (size_t)node * kNodeF4;
    const float4 lx = __ldg(g + h), ly = __ldg(g + 2 + h), lz = __ldg(g + 4 + h);
    const float4 hx = __ldg(g + 6 + h), hy = __ldg(g + 8 + h), hz = __ldg(g + 10 + h);
    const int4 r01 = __ldg(reinterpret_cast<const int4*>(g + 12) + 2 * h), r23 = __ldg(reinterpret_cast<const int4*>(g + 12) + 2 * h + 1);
    rx[0] = r01.x, ry[0] = r01.y, rx[1] = r01.z, ry[1] = r01.w, rx[2] = r23.x, ry[2] = r23.y, rx[3] = r23.z, ry[3] = r23.w;
    b[0] = box_bound_rn(qx, qy, qz, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x);
    b[1] = box_bound_rn(qx, qy, qz, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y);
    b[2] = box_bound_rn(qx, qy, qz, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z);
    b[3] = box_bound_rn(qx, qy, qz, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (ry[c] < 0 && rx[c] == skip) b[c] = inf;  // only the start node can have this child; ids are unique
  }
  // ---- leaf children, nearest first, each re-tested against the shrinking best distance
  bool lf[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) lf[c] = ry[c] > 0;
  for (;;) {
    float lb = inf;
    int lc = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (lf[c] && b[c] <= rs.worst() && b[c] < lb) {
        lb = b[c];
        lc = c;
      }
    const float ob = __shfl_xor_sync(kFull, lb, 1);
    const bool any_leaf = fminf(lb, ob) < inf;
    if (!__any_sync(kFull, any_leaf)) break;
    const bool mine = lb < ob || (lb == ob && h == 0);
    int st = sel4(rx, lc), cnt = sel4(ry, lc);
    const int ost = __shfl_xor_sync(kFull, st, 1), ocnt = __shfl_xor_sync(kFull, cnt, 1);
    if (!mine) {
      st = ost;
      cnt = ocnt;
    } else if (lb < inf) {
#pragma unroll
      for (int c = 0; c < 4; ++c) lf[c] = lf[c] && c != lc;
    }
    rs.scan(any_leaf, ix.spts, st, cnt, qx, qy, qz, h);
  }
  // ---- internal children: enter the nearest, keep the others
  float nb = inf;
  int nc = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (ry[c] < 0 && b[c] <= rs.worst() && b[c] < nb) {
      nb = b[c];
      nc = c;
    }
  const float onb = __shfl_xor_sync(kFull, nb, 1);
  const bool take = nb < onb || (nb == onb && h == 0);  // this lane holds the child to enter (if any)
  int next = sel4(rx, nc);
  const int onext = __shfl_xor_sync(kFull, next, 1);
  if (!take) next = onext;
  const bool descend = fminf(nb, onb) < inf;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (ry[c] < 0 && b[c] <= rs.worst() && !(take && descend && c == nc))
      stk[sp++] = ((unsigned long long)__float_as_uint(b[c]) << 32) | (unsigned)rx[c];
  // ---- next node: the child just chosen, else the nearer of the two stack tops that still matters
  while (sp > 0 && __uint_as_float((unsigned)(stk[sp - 1] >> 32)) > rs.worst()) --sp;
  const unsigned long long top = (run && !descend && sp > 0) ? stk[sp - 1] : ~0ull;
  const unsigned long long otop = __shfl_xor_sync(kFull, top, 1);
  if (run) {
    if (descend) {
      node = (unsigned)next;
    } else {
      const unsigned long long pick = top < otop ? top : otop;
      if (pick == ~0ull) {
        run = false;
      } else {
        node = (unsigned)(pick & 0xffffffffull);
        if (top < otop || (top == otop && h == 0)) --sp;
      }
    }
  }
}

// Depth-first search below node `start` by a pair of lanes.  MUST be called by all 32 lanes;
// `active`, the query and rs are uniform per pair.
template <class RS>
__device__ __forceinline__ void nn1_traverse_pair(const IndexView& ix, bool active, float qx, float qy, float qz, RS& rs,
                                                  int start = 0, int skip = -1) {
  const int h = threadIdx.x & 1;
  unsigned long long stk[kPairStack];
  int sp = 0;
  bool run = active && ix.n > 0;
  unsigned node = (unsigned)start;
  while (__any_sync(kFull, run)) nn1_visit_pair(ix, run, qx, qy, qz, rs, node, skip, stk, sp, h);
}

}  // namespace ddlo
