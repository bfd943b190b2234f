// Exact k-nearest-neighbour traversal of the 8-wide box tree (device code).
//
// Replaces nanoflann's KDTreeSingleIndexAdaptor::findNeighbors / searchLevel + KNNResultSet
// (R/impl/nanoflann_impl.hpp:1365-1384, 1495-1566, 161-243) for the two searches the engine does:
// k-NN of every point of a cloud in itself (covariances) and 1-NN of a moved source point in the
// target (correspondences).
//
// Exactness.  The squared distance is evaluated exactly as nanoflann's L2_Simple_Adaptor does
// (nanoflann_impl.hpp:508-517): d2 = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)) in float, with
// explicit round-to-nearest intrinsics so that nvcc can not contract it into FMAs.  A box's lower
// bound uses the same expression tree on the per-axis gaps; IEEE subtraction, multiplication and
// addition are monotone, hence bound <= d2 in float for every point inside the box, and a subtree
// is skipped only if bound > current k-th distance.  Candidates are ranked by (d2, original
// index), so ties are broken by index and the result equals a brute-force scan bit for bit.
#pragma once

#include "common.cuh"

namespace ddlo {

__device__ __forceinline__ float sqdist3_rn(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float axis_gap(float q, float lo, float hi) {
  // q - lo (negative) below the box, q - hi (positive) above it, 0 inside; only the square is used
  return q < lo ? __fsub_rn(q, lo) : (q > hi ? __fsub_rn(q, hi) : 0.0f);
}

__device__ __forceinline__ float box_bound_rn(float qx, float qy, float qz, float lx, float ly, float lz, float hx, float hy,
                                              float hz) {
  const float gx = axis_gap(qx, lx, hx), gy = axis_gap(qy, ly, hy), gz = axis_gap(qz, lz, hz);
  return __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
}

// bounds of the 8 sibling boxes of one group
__device__ __forceinline__ void group_bounds(const float4* __restrict__ g, float qx, float qy, float qz, float b[8]) {
  const float4 lx0 = __ldg(g + 0), lx1 = __ldg(g + 1), ly0 = __ldg(g + 2), ly1 = __ldg(g + 3), lz0 = __ldg(g + 4), lz1 = __ldg(g + 5);
  const float4 hx0 = __ldg(g + 6), hx1 = __ldg(g + 7), hy0 = __ldg(g + 8), hy1 = __ldg(g + 9), hz0 = __ldg(g + 10), hz1 = __ldg(g + 11);
  b[0] = box_bound_rn(qx, qy, qz, lx0.x, ly0.x, lz0.x, hx0.x, hy0.x, hz0.x);
  b[1] = box_bound_rn(qx, qy, qz, lx0.y, ly0.y, lz0.y, hx0.y, hy0.y, hz0.y);
  b[2] = box_bound_rn(qx, qy, qz, lx0.z, ly0.z, lz0.z, hx0.z, hy0.z, hz0.z);
  b[3] = box_bound_rn(qx, qy, qz, lx0.w, ly0.w, lz0.w, hx0.w, hy0.w, hz0.w);
  b[4] = box_bound_rn(qx, qy, qz, lx1.x, ly1.x, lz1.x, hx1.x, hy1.x, hz1.x);
  b[5] = box_bound_rn(qx, qy, qz, lx1.y, ly1.y, lz1.y, hx1.y, hy1.y, hz1.y);
  b[6] = box_bound_rn(qx, qy, qz, lx1.z, ly1.z, lz1.z, hx1.z, hy1.z, hz1.z);
  b[7] = box_bound_rn(qx, qy, qz, lx1.w, ly1.w, lz1.w, hx1.w, hy1.w, hz1.w);
}

// ---- result sets ------------------------------------------------------------------------------
// Both rank candidates by (d2, original index).  `worst()` is the current k-th distance, FLT_MAX
// until k candidates are known; like nanoflann (KNNResultSet::init, :181-187) a candidate whose
// distance is not below FLT_MAX is never admitted.

struct Best1 {
  float d = FLT_MAX;
  int idx = -1;   // original index
  int pos = -1;   // position in spts
  __device__ __forceinline__ float worst() const { return d; }
  __device__ __forceinline__ void offer(float dist, int oidx, int p) {
    if (dist < d || (dist == d && oidx < idx)) {
      d = dist;
      idx = oidx;
      pos = p;
    }
  }
};

// k entries per thread in shared memory, entry j of thread t at [j * stride + t] (conflict free)
struct TopKShared {
  float* d;
  int* idx;
  int k;
  int stride;
  __device__ __forceinline__ void init(float* d_, int* idx_, int k_, int stride_, int t) {
    d = d_ + t;
    idx = idx_ + t;
    k = k_;
    stride = stride_;
    for (int j = 0; j < k; ++j) {
      d[j * stride] = FLT_MAX;
      idx[j * stride] = -1;
    }
  }
  __device__ __forceinline__ float worst() const { return d[(k - 1) * stride]; }
  __device__ __forceinline__ void offer(float dist, int oidx, int /*pos*/) {
    const float wd = d[(k - 1) * stride];
    if (!(dist < wd || (dist == wd && oidx < idx[(k - 1) * stride]))) return;
    int j = k - 1;
    while (j > 0) {
      const float pd = d[(j - 1) * stride];
      if (pd > dist || (pd == dist && idx[(j - 1) * stride] > oidx)) {
        d[j * stride] = pd;
        idx[j * stride] = idx[(j - 1) * stride];
        --j;
      } else {
        break;
      }
    }
    d[j * stride] = dist;
    idx[j * stride] = oidx;
  }
};

__device__ __forceinline__ float pick8(const float b[8], int c) {
  float r = b[0];
#pragma unroll
  for (int s = 1; s < 8; ++s) r = (c == s) ? b[s] : r;
  return r;
}

template <class RS>
__device__ __forceinline__ void scan_leaf(const float4* __restrict__ spts, int leaf, float qx, float qy, float qz, RS& rs) {
  const float4* p = spts + (size_t)leaf * kLeaf;
#pragma unroll
  for (int j = 0; j < kLeaf; ++j) {
    const float4 v = __ldg(p + j);
    const float dist = sqdist3_rn(qx, qy, qz, v.x, v.y, v.z);  // padding points are +inf -> dist = +inf
    rs.offer(dist, __float_as_int(v.w), leaf * kLeaf + j);
  }
}

// Depth-first traversal, nearest child first, without a stack in memory: the pending children of
// the node being expanded on each level are one byte of `pend` (8 levels x 8 children).
template <class RS>
__device__ __forceinline__ void knn_traverse(const IndexView& ix, float qx, float qy, float qz, RS& rs) {
  if (ix.n <= 0) return;
  const int leaf_lvl = ix.nlev - 1;
  unsigned long long pend = 0ull;
  int lvl = 0;
  unsigned group = 0;  // index of the group being expanded on level `lvl` (= its parent's node id)
  for (;;) {
    float b[8];
    group_bounds(ix.box[lvl] + (size_t)group * 12, qx, qy, qz, b);
    bool descended = false;
    if (lvl == leaf_lvl) {
      // children are leaves: visit the nearest first, then every other one that still qualifies
      // (the bound is re-tested against the shrinking k-th distance right before each visit)
      const float w = rs.worst();
      unsigned m = 0;
      int c = 0;
      float bmin = FLT_MAX;
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (b[s] <= w) {
          m |= 1u << s;
          if (b[s] < bmin) {
            bmin = b[s];
            c = s;
          }
        }
      while (m) {
        m &= ~(1u << c);
        if (pick8(b, c) <= rs.worst()) scan_leaf(ix.spts, (int)(group * 8 + c), qx, qy, qz, rs);
        c = __ffs(m) - 1;
      }
    } else {
      const float w = rs.worst();
      unsigned m = 0;
      int cmin = -1;
      float bmin = 0.0f;
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (b[s] <= w) {
          m |= 1u << s;
          if (cmin < 0 || b[s] < bmin) {
            bmin = b[s];
            cmin = s;
          }
        }
      if (m) {
        m &= ~(1u << cmin);
        pend = (pend & ~(0xffull << (8 * lvl))) | ((unsigned long long)m << (8 * lvl));
        group = group * 8 + cmin;
        ++lvl;
        descended = true;
      }
    }
    if (descended) continue;
    // pop: climb until some level has a pending child, then descend into it
    for (;;) {
      if (lvl == 0) return;
      --lvl;
      group >>= 3;
      const unsigned m = (unsigned)((pend >> (8 * lvl)) & 0xffull);
      if (m) {
        const int c = __ffs(m) - 1;
        pend &= ~(1ull << (8 * lvl + c));
        group = group * 8 + c;
        ++lvl;
        break;
      }
    }
  }
}

}  // namespace ddlo
