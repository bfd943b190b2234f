// Exact k-nearest-neighbour search in the Morton-prefix octree (device code).
//
// Replaces nanoflann's KDTreeSingleIndexAdaptor::findNeighbors / searchLevel + KNNResultSet
// (R/impl/nanoflann_impl.hpp:1365-1384, 1495-1566, 161-243) for the two searches the engine does:
// k-NN of every point of a cloud in itself (covariances) and 1-NN of a moved source point in the
// target (correspondences).
//
// Mapping.  One query is served by a SUB-WARP of 8 lanes (4 queries per warp): a node has 8 child
// boxes, so lane s tests child s with one coalesced 32-byte load per box component, and a leaf's
// points are scanned 8 at a time.  All lanes of a sub-warp hold the same traversal state; the four
// sub-warps of a warp run in lock step and idle when their query is finished.  Compared with one
// thread per query this removes the divergence between neighbouring queries and multiplies the
// number of resident warps by eight, which is what hides the L2 latency of the pointer chase.
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

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSubLanes = 8;
constexpr int kIdxSentinel = 0x7fffffff;

__device__ __forceinline__ float sqdist3_rn(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float axis_gap(float q, float lo, float hi) {
  // |q - lo| below the box, |q - hi| above it, 0 inside (only the square is used).  Branch-free:
  // fl(lo - q) == -fl(q - lo) exactly, at most one of the two differences is positive, and an empty
  // slot (lo = +inf, hi = -inf) gives +inf.
  return fmaxf(fmaxf(__fsub_rn(lo, q), __fsub_rn(q, hi)), 0.0f);
}

__device__ __forceinline__ float box_bound_rn(float qx, float qy, float qz, float lx, float ly, float lz, float hx, float hy,
                                              float hz) {
  const float gx = axis_gap(qx, lx, hx), gy = axis_gap(qy, ly, hy), gz = axis_gap(qz, lz, hz);
  return __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
}

// (d, i) lexicographic order
__device__ __forceinline__ bool lex_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

// a < b for two candidate keys bits(d) << 32 | index (d a finite non-negative float, or the empty key (FLT_MAX, sentinel),
// or 0).  Such a key is also the bit pattern of a finite non-negative DOUBLE (its exponent field is the top 11 bits of
// bits(d) < 0x7f800000, never all ones), and finite non-negative doubles - zero and denormals included, double
// arithmetic never flushes them - order exactly like their bit patterns.  So the comparison is one DSETP instead of an
// ISETP pair.  NOT for the "no candidate" pattern ~0 (a NaN as a double): callers compare that one as an integer.
__device__ __forceinline__ bool key_less(unsigned long long a, unsigned long long b) {
  return __longlong_as_double((long long)a) < __longlong_as_double((long long)b);
}

struct Sub {
  int sl;          // lane inside the sub-warp, 0..7
  int base;        // first lane of the sub-warp inside the warp
  unsigned shift;  // bit offset of the sub-warp's byte in a warp ballot
};
__device__ __forceinline__ Sub make_sub() {
  const int lane = threadIdx.x & 31;
  Sub s;
  s.sl = lane & 7;
  s.base = lane & ~7;
  s.shift = (unsigned)(lane & ~7);
  return s;
}
__device__ __forceinline__ unsigned sub_ballot(bool p, const Sub& sb) { return (__ballot_sync(kFull, p) >> sb.shift) & 0xffu; }
// Slot (0..7) of the lane holding the smallest non-negative float among the lanes with `cand`, -1 if
// none.  Three butterfly steps on one packed key: the float's bits (which order like unsigned ints
// for v >= 0) with the 3 low mantissa bits replaced by the slot.  The choice only steers the order
// in which children are visited, so losing 3 mantissa bits of the bound is harmless; pruning always
// uses the exact bound.  (redux.sync with a sub-warp mask would serialise the four sub-warps.)
__device__ __forceinline__ int sub_argmin(bool cand, float v, const Sub& sb) {
  unsigned key = cand ? ((__float_as_uint(v) & 0xfffffff8u) | (unsigned)sb.sl) : 0xffffffffu;
#pragma unroll
  for (int o = 1; o < kSubLanes; o <<= 1) key = min(key, __shfl_xor_sync(kFull, key, o));
  return key == 0xffffffffu ? -1 : (int)(key & 7u);
}
// ---- result sets ------------------------------------------------------------------------------
// Both rank candidates by (d2, original index).  `worst()` is the current k-th distance, FLT_MAX
// until k candidates are known; like nanoflann (KNNResultSet::init, nanoflann_impl.hpp:181-187) a
// candidate whose distance is not below FLT_MAX is never admitted.  All members are uniform across
// the 8 lanes of a sub-warp.  Empty = (FLT_MAX, kIdxSentinel), the largest possible pair.

// k = 1: best candidate in registers.  It may be seeded with any real point of the cloud (its exact
// distance and index): a seed only shrinks the search, the result stays the exact (d, i) minimum.
struct Best1Sub {
  float d = FLT_MAX;
  int idx = kIdxSentinel;  // original index
  int pos = -1;            // position in spts (Morton order)
  __device__ __forceinline__ float worst() const { return d; }
  // (d, idx) as one 64-bit key: for d >= 0 the float's bits order like an unsigned int, so the
  // lexicographic minimum is the plain unsigned minimum of (bits(d) << 32 | idx)
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
  __device__ __forceinline__ void scan(bool doit, const float4* __restrict__ spts, int start, int count, float qx, float qy, float qz,
                                       const Sub& sb) {
    unsigned long long key = pack(d, idx);
    int lp = pos;
    if (doit) {
      for (int j = sb.sl; j < count; j += kSubLanes) {
        const float4 v = __ldg(spts + start + j);
        const float dist = sqdist3_rn(qx, qy, qz, v.x, v.y, v.z);
        const unsigned long long k2 = pack(dist, __float_as_int(v.w));
        if (dist < FLT_MAX && k2 < key) {
          key = k2;
          lp = start + j;
        }
      }
    }
    unsigned long long m = key;
#pragma unroll
    for (int o = 1; o < kSubLanes; o <<= 1) m = min(m, __shfl_xor_sync(kFull, m, o));
    const unsigned who = sub_ballot(key == m, sb);  // never empty; all eight when nothing improved
    pos = __shfl_sync(kFull, lp, sb.base + __ffs(who) - 1);
    d = __uint_as_float((unsigned)(m >> 32));
    idx = (int)(unsigned)(m & 0xffffffffull);
  }
};

// general k: k unsorted (d, i) pairs in shared memory plus the tracked maximum.  Empty slots hold
// the sentinel, so "replace the maximum" also fills the set.
struct TopKSub {
  float* sd;
  int* si;
  int k;
  float wd;
  int wi;
  int wslot;
  __device__ __forceinline__ void init(float* sd_, int* si_, int k_, const Sub& sb) {
    sd = sd_;
    si = si_;
    k = k_;
    for (int e = sb.sl; e < k; e += kSubLanes) {
      sd[e] = FLT_MAX;
      si[e] = kIdxSentinel;
    }
    wd = FLT_MAX;
    wi = kIdxSentinel;
    wslot = k - 1;
    __syncwarp();
  }
  __device__ __forceinline__ float worst() const { return wd; }
  // maximum over the k slots; equal pairs (only sentinels can be equal) resolved by slot so that
  // every lane agrees.  Contains warp-wide shuffles: must be executed by all 32 lanes.
  __device__ __forceinline__ void find_max(const Sub& sb) {
    float md = -1.0f;
    int mi = -1, ms = -1;
    for (int e = sb.sl; e < k; e += kSubLanes) {
      const float ed = sd[e];
      const int ei = si[e];
      if (ms < 0 || lex_less(md, mi, ed, ei) || (ed == md && ei == mi && e > ms)) {
        md = ed;
        mi = ei;
        ms = e;
      }
    }
#pragma unroll
    for (int o = 1; o < kSubLanes; o <<= 1) {
      const float od = __shfl_xor_sync(kFull, md, o);
      const int oi = __shfl_xor_sync(kFull, mi, o);
      const int os = __shfl_xor_sync(kFull, ms, o);
      if (os >= 0 && (ms < 0 || lex_less(md, mi, od, oi) || (od == md && oi == mi && os > ms))) {
        md = od;
        mi = oi;
        ms = os;
      }
    }
    wd = md;
    wi = mi;
    wslot = ms;
  }
  __device__ __forceinline__ void scan(bool doit, const float4* __restrict__ spts, int start, int count, float qx, float qy, float qz,
                                       const Sub& sb) {
    for (int r = 0; __any_sync(kFull, doit && r < count); r += kSubLanes) {
      const int j = r + sb.sl;
      const bool have = doit && j < count;
      float dist = FLT_MAX;
      int oi = kIdxSentinel;
      if (have) {
        const float4 v = __ldg(spts + start + j);
        dist = sqdist3_rn(qx, qy, qz, v.x, v.y, v.z);
        oi = __float_as_int(v.w);
      }
      unsigned cb = sub_ballot(have && dist < FLT_MAX && lex_less(dist, oi, wd, wi), sb);
      while (__any_sync(kFull, cb != 0u)) {
        const bool act = cb != 0u;
        const int c = act ? __ffs(cb) - 1 : 0;
        cb &= cb - 1u;
        const float dc = __shfl_sync(kFull, dist, sb.base + c);
        const int ic = __shfl_sync(kFull, oi, sb.base + c);
        const bool upd = act && lex_less(dc, ic, wd, wi);  // the bar may have dropped since the ballot
        if (upd && sb.sl == 0) {
          sd[wslot] = dc;
          si[wslot] = ic;
        }
        __syncwarp();
        find_max(sb);
      }
    }
  }
  // ascending (d, i) order into one output row; sentinels become (-1, +inf)
  __device__ __forceinline__ void write_sorted(int* __restrict__ idx_out, float* __restrict__ d_out, const Sub& sb) const {
    for (int e = sb.sl; e < k; e += kSubLanes) {
      const float ed = sd[e];
      const int ei = si[e];
      int rank = 0;
      for (int f = 0; f < k; ++f) rank += (lex_less(sd[f], si[f], ed, ei) || (sd[f] == ed && si[f] == ei && f < e)) ? 1 : 0;
      const bool empty = ei == kIdxSentinel;
      idx_out[rank] = empty ? -1 : ei;
      if (d_out) d_out[rank] = empty ? __int_as_float(0x7f800000) : ed;
    }
  }
};

// k <= 8*R: the candidates sorted in registers, R per lane, position = lane * R + r, so lane 0 holds
// the R smallest.  A candidate is one 64-bit key, bits(d) << 32 | index: for d >= 0 the unsigned
// order of the keys is the lexicographic (d, i) order.  Inserting one candidate is a shift of the
// tail by one position: every slot takes its left neighbour if that neighbour is greater than the
// candidate, the candidate itself if only the slot's own content is greater, else keeps its
// content.  The bar is the entry at position k-1, re-broadcast after every insertion, optionally
// capped from outside by any proven upper bound of the k-th distance (`cap`: e.g. the largest
// distance to k arbitrary points of the cloud), which keeps far candidates out while the list fills.
template <int R>
struct TopKRegSub {
  unsigned long long e[R];
  unsigned long long bar;  // candidates must be < bar
  unsigned long long capk; // (cap, sentinel index): admits every index at distance == cap
  int k;
  int skip_lo = 0, skip_hi = 0;  // Morton positions already in the list (prefill): scans leave them out
  static __device__ __forceinline__ unsigned long long pack(float dist, int oi) {
    return ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)oi;
  }
  __device__ __forceinline__ void init(int k_, float cap = FLT_MAX) {
    k = k_;
#pragma unroll
    for (int r = 0; r < R; ++r) e[r] = pack(FLT_MAX, kIdxSentinel);
    capk = pack(cap, kIdxSentinel);
    bar = capk;
  }
  __device__ __forceinline__ float worst() const { return __uint_as_float((unsigned)(bar >> 32)); }
  // the bar from the entry at position k-1 (all 32 lanes)
  __device__ __forceinline__ void refresh_bar(const Sub& sb) {
    const int kr = (k - 1) % R, kl = sb.base + (k - 1) / R;
    unsigned long long kth = e[0];
#pragma unroll
    for (int r = 1; r < R; ++r) kth = kr == r ? e[r] : kth;
    kth = __shfl_sync(kFull, kth, kl);
    bar = min(kth, capk);
  }
  // Start from 8R real points of the cloud, the Morton positions [w0, w0 + 8R), instead of an empty list: their
  // keys are ranked against each other (every lane compares its R keys with the R keys of each other lane) and
  // moved to their sorted positions through `scratch` (8R words of shared memory owned by the sub-warp).  This
  // replaces the first k insertions by one pass of compares, and the bar starts at the k-th distance inside the
  // window.  Scans must skip these positions (skip_lo / skip_hi).  All 32 lanes; `active` uniform per sub-warp.
  __device__ __forceinline__ void prefill(bool active, const float4* __restrict__ spts, int w0, float qx, float qy, float qz,
                                          unsigned long long* __restrict__ scratch, const Sub& sb) {
    unsigned long long key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      key[r] = pack(FLT_MAX, kIdxSentinel - (sb.sl * R + r));  // (distinct dummies for idle sub-warps)
      if (active) {
        const float4 v = __ldg(spts + w0 + sb.sl * R + r);
        // (a distance that overflowed - or is NaN, in a cloud the checked entry points reject afterwards - is entered as
        // FLT_MAX: every key stays a finite double for key_less)
        key[r] = pack(fminf(sqdist3_rn(qx, qy, qz, v.x, v.y, v.z), FLT_MAX), __float_as_int(v.w));
      }
    }
    int rank[R];
    auto lt = [](unsigned long long x, unsigned long long y) { return key_less(x, y); };
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rank[r] = 0;
#pragma unroll
      for (int q = 0; q < R; ++q) rank[r] += lt(key[q], key[r]) ? 1 : 0;
    }
#pragma unroll
    for (int d = 1; d < kSubLanes; ++d) {
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const unsigned long long o = __shfl_sync(kFull, key[q], sb.base + ((sb.sl + d) & (kSubLanes - 1)));
#pragma unroll
        for (int r = 0; r < R; ++r) rank[r] += lt(o, key[r]) ? 1 : 0;
      }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) scratch[rank[r]] = key[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) e[r] = scratch[sb.sl * R + r];
    __syncwarp();
    if (active) {
      skip_lo = w0;
      skip_hi = w0 + kSubLanes * R;
    }
    refresh_bar(sb);
  }
  // all 32 lanes; `c` and `upd` uniform per sub-warp.  Puts c at its sorted position (the last entry falls off);
  // the bar is NOT refreshed: a candidate admitted against a slightly stale bar lands behind position k-1 and
  // changes nothing among the k nearest, so one refresh per scan step is enough.
  __device__ __forceinline__ void insert_only(bool upd, unsigned long long c, const Sub& sb) {
    unsigned long long left = __shfl_up_sync(kFull, e[R - 1], 1);
    if (sb.sl == 0) left = 0ull;  // nothing to the left of position 0
    if (upd) {
      // g[r]: the content of slot r - 1 (the left neighbour's last entry for r = 0) is greater than c.  From the tail: a
      // slot greater than c takes its left neighbour if that one is greater too, else c itself.  The comparisons are made
      // on the keys READ AS DOUBLES (key_less): one DSETP instead of the ISETP pair of a 64-bit integer comparison, and the
      // selects become predicated moves.  Loop body 52 -> 30 instructions; covariances of the C2 step 0.137 -> 0.121 ms.
      bool g[R + 1];
      g[0] = key_less(c, left);
#pragma unroll
      for (int r = 0; r < R; ++r) g[r + 1] = key_less(c, e[r]);
#pragma unroll
      for (int r = R - 1; r >= 0; --r) {
        const unsigned long long lv = r == 0 ? left : e[r == 0 ? 0 : r - 1];
        if (g[r + 1]) e[r] = g[r] ? lv : c;
      }
    }
  }
  __device__ __forceinline__ void insert(bool upd, unsigned long long c, const Sub& sb) {
    insert_only(upd, c, sb);
    refresh_bar(sb);
  }
  __device__ __forceinline__ void scan(bool doit, const float4* __restrict__ spts, int start, int count, float qx, float qy, float qz,
                                       const Sub& sb) {
    for (int r = 0; __any_sync(kFull, doit && r < count); r += kSubLanes) {
      const int j = r + sb.sl;
      unsigned long long key = ~0ull;
      if (doit && j < count && !(start + j >= skip_lo && start + j < skip_hi)) {
        const float4 v = __ldg(spts + start + j);
        const float dist = sqdist3_rn(qx, qy, qz, v.x, v.y, v.z);
        if (dist < FLT_MAX) key = pack(dist, __float_as_int(v.w));
      }
      unsigned cb = sub_ballot(key < bar, sb);
      if (__any_sync(kFull, cb != 0u)) {
        do {
          const bool act = cb != 0u;
          const int c = act ? __ffs(cb) - 1 : 0;
          cb &= cb - 1u;
          const unsigned long long ck = __shfl_sync(kFull, key, sb.base + c);
          insert_only(act, ck, sb);
        } while (__any_sync(kFull, cb != 0u));
        refresh_bar(sb);
      }
    }
  }
  __device__ __forceinline__ void write_sorted(int* __restrict__ idx_out, float* __restrict__ d_out, const Sub& sb) const {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int p = sb.sl * R + r;
      if (p < k) {
        const int oi = (int)(unsigned)(e[r] & 0xffffffffull);
        const bool empty = oi == kIdxSentinel;
        idx_out[p] = empty ? -1 : oi;
        if (d_out) d_out[p] = empty ? __int_as_float(0x7f800000) : __uint_as_float((unsigned)(e[r] >> 32));
      }
    }
  }
};

// ---- starting deep in the tree ------------------------------------------------------------------
// A node is a cube of the Morton lattice.  If the ball around the query with the current k-th
// distance as squared radius lies inside that cube, no point outside the node's subtree can enter
// the result, so the search may start (or stop climbing) there.  The test is made on the lattice in
// float, with the radius inflated by 1e-5 relative and 1e-3 lattice units absolute: far more than
// the rounding of u = (p - lo) * scale (a few 1e-4 lattice units at u <= 1023) and of the float
// distances, so every point within the ball provably got a code inside the cube.  Cubes on the
// lattice border are open-ended because out-of-range coordinates were clamped into them.
struct Ball {  // the query and its radius on the lattice
  float u[3];
  float ur;
  bool finite;
};
__device__ __forceinline__ Ball make_ball(const IndexView& ix, float qx, float qy, float qz, float worst) {
  Ball b;
  b.finite = worst < FLT_MAX;
  const float4 lat = __ldg(reinterpret_cast<const float4*>(ix.lattice));
  b.ur = sqrtf(worst) * 1.00001f * lat.w + 1e-3f;
  b.u[0] = (qx - lat.x) * lat.w;
  b.u[1] = (qy - lat.y) * lat.w;
  b.u[2] = (qz - lat.z) * lat.w;
  return b;
}
__device__ __forceinline__ bool ball_in_cell(const Ball& b, const int4 m) {
  if (m.y == 0) return true;  // the root holds everything
  if (!b.finite) return false;
  const float w = (float)(1 << (kMortonLevels - m.y));
  bool ok = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float c0 = (float)((m.z >> (10 * a)) & 1023);
    ok = ok && (c0 == 0.0f || b.u[a] - b.ur >= c0) && (c0 + w >= 1024.0f || b.u[a] + b.ur < c0 + w);
  }
  return ok;
}

// the deepest node around a known point of the cloud (original index j) whose cube holds the ball
__device__ __forceinline__ int start_node_for(const IndexView& ix, int j, float qx, float qy, float qz, float worst) {
  const Ball b = make_ball(ix, qx, qy, qz, worst);
  int node = __ldg(ix.node_of_point + j);
  for (;;) {
    const int4 m = __ldg(ix.meta + node);
    if (ball_in_cell(b, m)) return node;
    node = m.x;
  }
}

// Depth-first traversal by one sub-warp.  At a node: lane s bounds child s; children whose bound
// can still beat the k-th distance ("<=", so that an equal-distance point with a smaller index is
// not missed) are handled leaves first, nearest first, each re-tested against the shrinking bar
// right before its scan; of the internal ones the nearest is entered directly and the others are
// pushed, with their bound, on the sub-warp's stack in shared memory, so that a stale entry is
// dropped on pop without touching global memory.
// MUST be called by all 32 lanes of a warp; `active` is uniform per sub-warp; `stack` has
// kStackDepth entries per sub-warp.
// `start` is the node to search below (0 = the whole cloud); `skip` (or -1) is a child NODE of
// `start` that has been searched already and is left out.
template <class RS>
__device__ __forceinline__ void knn_traverse_sub(const IndexView& ix, bool active, float qx, float qy, float qz, RS& rs,
                                                 unsigned long long* __restrict__ stack, const Sub& sb, int start = 0, int skip = -1) {
  bool run = active && ix.n > 0;
  int sp = 0;
  unsigned node = (unsigned)start;
  const float inf = __int_as_float(0x7f800000);
  while (__any_sync(kFull, run)) {
    __syncwarp();
    float b = inf;
    int2 ref = make_int2(0, 0);
    if (run) {
      DDLO_CHECK_INDEX(node, reinterpret_cast<const int*>(ix.lattice)[5], "knn_traverse_sub: node");
      const float* g = reinterpret_cast<const float*>(ix.nodes + (size_t)node * kNodeF4);
      const float lx = __ldg(g + sb.sl), ly = __ldg(g + 8 + sb.sl), lz = __ldg(g + 16 + sb.sl);
      const float hx = __ldg(g + 24 + sb.sl), hy = __ldg(g + 32 + sb.sl), hz = __ldg(g + 40 + sb.sl);
      ref = __ldg(reinterpret_cast<const int2*>(g + 48) + sb.sl);
      b = box_bound_rn(qx, qy, qz, lx, ly, lz, hx, hy, hz);
      if (ref.y < 0 && ref.x == skip) b = inf;  // only `start` can have this child; ids are unique
    }
    const bool qual = run && b <= rs.worst();
    // ---- leaf children, nearest first
    unsigned ml = sub_ballot(qual && ref.y > 0, sb);
    while (__any_sync(kFull, ml != 0u)) {
      const int am = sub_argmin(((ml >> sb.sl) & 1u) != 0u, b, sb);
      const int mc = am < 0 ? 0 : am;
      const float mb = __shfl_sync(kFull, b, sb.base + mc);
      const bool had = ml != 0u;
      ml &= ~(1u << mc);
      const int st = __shfl_sync(kFull, ref.x, sb.base + mc);
      const int cnt = __shfl_sync(kFull, ref.y, sb.base + mc);
      if (had) DDLO_CHECK_INDEX(st + cnt - 1, ix.n, "knn_traverse_sub: leaf range");
      rs.scan(had && mb <= rs.worst(), ix.spts, st, cnt, qx, qy, qz, sb);
    }
    // ---- internal children
    const bool qi = qual && ref.y < 0 && b <= rs.worst();
    const unsigned mi = sub_ballot(qi, sb);
    const int an = sub_argmin(qi, b, sb);
    const int nc = an < 0 ? 0 : an;
    const unsigned rest = mi & ~(1u << nc);
    if (qi && sb.sl != nc) {
      const int at = sp + __popc(rest & ((1u << sb.sl) - 1u));
      stack[at] = ((unsigned long long)__float_as_uint(b) << 32) | (unsigned)ref.x;
    }
    const int next = __shfl_sync(kFull, ref.x, sb.base + nc);
    __syncwarp();
    if (run) {
      if (mi) {
        sp += __popc(rest);
        node = (unsigned)next;
      } else {
        // pop the next entry that can still matter (bounds are >= 0, so their bits order like floats)
        run = false;
        while (sp > 0) {
          const unsigned long long e = stack[--sp];
          if (__uint_as_float((unsigned)(e >> 32)) <= rs.worst()) {
            node = (unsigned)(e & 0xffffffffull);
            run = true;
            break;
          }
        }
      }
    }
  }
}

}  // namespace ddlo
