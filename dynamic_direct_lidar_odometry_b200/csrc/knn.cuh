// Exact k-nearest-neighbour traversal of the Morton-prefix octree (device code).
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
__device__ __forceinline__ void scan_leaf(const float4* __restrict__ spts, int start, int count, float qx, float qy, float qz, RS& rs) {
  const float4* p = spts + start;
#pragma unroll 4
  for (int j = 0; j < count; ++j) {
    const float4 v = __ldg(p + j);
    rs.offer(sqdist3_rn(qx, qy, qz, v.x, v.y, v.z), __float_as_int(v.w), start + j);
  }
}

// Depth-first traversal.  At a node: the child boxes that can still hold a better candidate are
// found (bound <= current k-th distance; "<=" so that an equal-distance point with a smaller index
// is not missed); leaf children are scanned on the spot, which tightens the k-th distance before
// the internal children are looked at; of those the nearest is entered directly and the others go
// on a small per-thread stack together with their bound, so that a stale entry is discarded on
// pop without touching memory.
template <class RS>
__device__ __forceinline__ void knn_traverse(const IndexView& ix, float qx, float qy, float qz, RS& rs) {
  if (ix.n <= 0) return;
  unsigned long long stack[kStackDepth];  // (bound bits << 32) | node id; bounds are >= 0 so the bits order like floats
  int sp = 0;
  unsigned node = 0;
  for (;;) {
    const float4* g = ix.nodes + (size_t)node * kNodeF4;
    float b[8];
    group_bounds(g, qx, qy, qz, b);
    const int2* refs = reinterpret_cast<const int2*>(g + 12);
    const float w = rs.worst();
    unsigned m = 0;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      if (b[s] <= w) m |= 1u << s;
    // leaves first
    unsigned inner = 0;
    for (unsigned mm = m; mm;) {
      const int s = __ffs(mm) - 1;
      mm &= mm - 1;
      const int2 r = __ldg(refs + s);
      if (r.y > 0) {
        if (pick8(b, s) <= rs.worst()) scan_leaf(ix.spts, r.x, r.y, qx, qy, qz, rs);
      } else {
        inner |= 1u << s;
      }
    }
    // internal children: nearest is entered now, the rest are stacked
    float best_b = 0.0f;
    int best_n = -1;
    const float w2 = rs.worst();
    for (; inner;) {
      const int s = __ffs(inner) - 1;
      inner &= inner - 1;
      const float bs = pick8(b, s);
      if (!(bs <= w2)) continue;
      const int child = __ldg(refs + s).x;
      float push_b;
      int push_n;
      if (best_n < 0 || bs < best_b) {
        push_b = best_b, push_n = best_n;
        best_b = bs, best_n = child;
      } else {
        push_b = bs, push_n = child;
      }
      if (push_n >= 0 && sp < kStackDepth) stack[sp++] = ((unsigned long long)__float_as_uint(push_b) << 32) | (unsigned)push_n;
    }
    if (best_n >= 0) {
      node = (unsigned)best_n;
      continue;
    }
    // pop the next entry that can still matter
    for (;;) {
      if (sp == 0) return;
      const unsigned long long e = stack[--sp];
      if (__uint_as_float((unsigned)(e >> 32)) <= rs.worst()) {
        node = (unsigned)(e & 0xffffffffull);
        break;
      }
    }
  }
}

}  // namespace ddlo
