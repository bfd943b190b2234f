// Bounding box + Morton keys + radix sort of a scan-sized cloud in ONE kernel: a thread-block cluster
// of 16 CTAs holds all (key, index) pairs in distributed shared memory and sorts them there.
//
// For the index build of the reference's kd-tree replacement (index.cu) the generic path is
// k_bounds -> k_morton -> cub::DeviceRadixSort (histogram, scan, 4 onesweep passes): ten launches and
// about 65 us for a 64x1024 scan, because a 65k-key sort is eight tiles of a library kernel tuned for
// hundreds of millions of keys.  A scan (<= 131072 points) fits the shared memory of 16 SMs, so here
//   phase 0  every CTA reduces the box of its slice; the 16 partial boxes are exchanged through DSMEM
//   phase 1  30-bit Morton keys on the lattice of that box (the arithmetic of k_morton, bit for bit)
//   phase 2  LSD radix sort, six passes of 5 bits.  Per pass every thread counts the digits of its own
//            contiguous chunk (private counters, no atomics), the counts are scanned per digit over the
//            threads of the CTA, and the CTA regroups its slice by digit in a staging buffer (chunks are
//            walked in order: stable).  After one cluster barrier every CTA reads the others' 32 digit
//            totals through DSMEM; a digit's run in the staging buffer is contiguous in the global order
//            too, so the exchange is consecutive threads storing consecutive 8-byte {key, value} elements
//            into the destination CTA's slice (st.shared::cluster, coalesced) - scattered 4-byte DSMEM
//            stores, one per key and value, were measured at 7 us per pass.  The output is identical to
//            the generic path's: same keys, same permutation.
//   phase 3  sorted keys / values to global memory, coalesced
// Two cluster barriers per pass, no global-memory traffic in between.  Clusters of 16 are non-portable
// (cudaFuncAttributeNonPortableClusterSizeAllowed); if the launch is refused the caller uses the generic path.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ddlo {

constexpr int kCsCtas = 16;
#ifndef DDLO_CS_BITS
#define DDLO_CS_BITS 5
#endif
constexpr int kCsBits = DDLO_CS_BITS;  // 5: six passes over 32 digits; 6: five passes over 64 digits (measured: index 95 vs 97 us, no gain)
constexpr int kCsBins = 1 << kCsBits;
constexpr int kCsPasses = (30 + kCsBits - 1) / kCsBits;
static_assert(kCsBins == 32 || kCsBins == 64, "a warp scans the digit totals: one or two digits per lane");

// exclusive scan of kCsBins values held by warp 0 (value of digit lane, and of digit lane + 32 when there are 64)
__device__ __forceinline__ void cs_scan_bins(unsigned lo_v, unsigned hi_v, int lane, unsigned& lo_ex, unsigned& hi_ex) {
  unsigned a = lo_v, b = hi_v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned ua = __shfl_up_sync(0xffffffffu, a, o), ub = __shfl_up_sync(0xffffffffu, b, o);
    if (lane >= o) a += ua, b += ub;
  }
  const unsigned total_lo = __shfl_sync(0xffffffffu, a, 31);
  lo_ex = a - lo_v;
  hi_ex = total_lo + b - hi_v;
}

__device__ __forceinline__ unsigned cs_spread10(unsigned v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

template <int kCsThreads, int IPT>
struct CsLayout {
  static constexpr int kCsWarps = kCsThreads / 32;
  static constexpr int kCap = kCsThreads * IPT;       // elements per CTA
  static constexpr int kStride = IPT | 1;             // odd chunk stride (in elements): conflict-free chunk walks
  static constexpr int kPhys = kCsThreads * kStride;  // physical elements of the chunked buffer A
  static constexpr size_t kBytes = (size_t)(kPhys + kCap) * 8                       // A (chunked, padded) + B (staging), uint2 {key, value}
                                   + (size_t)kCsBins * kCsThreads * 2               // cnt
                                   + (size_t)kCsBins * 4 * (3 + kCsCtas) + 8 * 4 + kCsWarps * 8 * 4;
  static __device__ __forceinline__ int phys(int i) { return (i / IPT) * kStride + (i % IPT); }
};

// lattice: {lo.x, lo.y, lo.z, scale, (int) non-finite count, (int) node count [written later]} as in index.cu
template <int kCsThreads, int IPT>
__global__ void __launch_bounds__(kCsThreads, 1) k_morton_sort_cluster(const float4* __restrict__ pts, int n, unsigned* __restrict__ keys_out,
                                                                         int* __restrict__ vals_out, float* __restrict__ lattice) {
  using L = CsLayout<kCsThreads, IPT>;
  constexpr int kCsWarps = L::kCsWarps;
  constexpr int kPerLane = kCsThreads / 32;        // counters of one digit a lane scans
  constexpr int kBinsPerWarp = (kCsBins + kCsWarps - 1) / kCsWarps;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  extern __shared__ __align__(16) unsigned char cs_smem[];
  uint2* A = reinterpret_cast<uint2*>(cs_smem);  // [kPhys]  this CTA's slice of the array, thread chunks padded to an odd stride
  uint2* B = A + L::kPhys;                       // [kCap]   the slice regrouped by digit (staging for the exchange)
  unsigned short* cnt = reinterpret_cast<unsigned short*>(B + L::kCap);  // [kCsBins][kCsThreads]
  unsigned* tot = reinterpret_cast<unsigned*>(cnt + kCsBins * kCsThreads);  // [kCsBins] this CTA's digit totals
  unsigned* lbase = tot + kCsBins;                                           // [kCsBins] start of the digit's run in B
  unsigned* base = lbase + kCsBins;                                          // [kCsBins] global start of (digit, this CTA)
  unsigned* all = base + kCsBins;                                            // [kCsBins][kCsCtas]
  float* cbox = reinterpret_cast<float*>(all + kCsBins * kCsCtas);           // [8] this CTA's box: lo3, hi3, bad
  float* wred = cbox + 8;                                                    // [kCsWarps][8]
  const int first = rank * L::kCap;  // global index of this CTA's first element

  // ---- phase 0: bounding box of the finite points (k_bounds)
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float bad = 0.0f;
  for (int i = t; i < L::kCap; i += kCsThreads) {
    const int gi = first + i;
    if (gi < n) {
      const float4 p = __ldg(pts + gi);
      if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
        bad = 1.0f;
      } else {
        lo[0] = fminf(lo[0], p.x), lo[1] = fminf(lo[1], p.y), lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x), hi[1] = fmaxf(hi[1], p.y), hi[2] = fmaxf(hi[2], p.z);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    bad = fmaxf(bad, __shfl_xor_sync(0xffffffffu, bad, o));
  }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      wred[warp * 8 + a] = lo[a];
      wred[warp * 8 + 3 + a] = hi[a];
    }
    wred[warp * 8 + 6] = bad;
  }
  __syncthreads();
  if (t < 7) {
    float v = wred[t];
    for (int w = 1; w < kCsWarps; ++w) v = t < 3 ? fminf(v, wred[w * 8 + t]) : fmaxf(v, wred[w * 8 + t]);
    cbox[t] = v;
  }
  cluster.sync();
  // every CTA folds the 16 partial boxes itself (through DSMEM) and gets the same result
  __shared__ float fbox[8];
  if (t < 7) {
    float v = cbox[t];
    for (int c = 0; c < kCsCtas; ++c) {
      const float o = cluster.map_shared_rank(cbox, c)[t];
      v = t < 3 ? fminf(v, o) : fmaxf(v, o);
    }
    fbox[t] = v;
  }
  __syncthreads();
  const float lx = fbox[0], ly = fbox[1], lz = fbox[2];
  const float ex = fbox[3] - lx, ey = fbox[4] - ly, ez = fbox[5] - lz;
  const float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-30f));
  const float scale = 1023.0f / ext;  // one isotropic lattice: cells stay cubes (k_morton)
  if (rank == 0 && t == 0) {
    lattice[0] = lx, lattice[1] = ly, lattice[2] = lz, lattice[3] = scale;
    reinterpret_cast<unsigned*>(lattice)[4] = fbox[6] != 0.0f ? 1u : 0u;
  }

  // ---- phase 1: Morton keys (k_morton); padding sorts behind every real point
  for (int i = t; i < L::kCap; i += kCsThreads) {
    const int gi = first + i;
    unsigned key = 0xffffffffu;
    if (gi < n) {
      const float4 p = __ldg(pts + gi);
      const unsigned cx = (unsigned)fminf(fmaxf((p.x - lx) * scale, 0.0f), 1023.0f);
      const unsigned cy = (unsigned)fminf(fmaxf((p.y - ly) * scale, 0.0f), 1023.0f);
      const unsigned cz = (unsigned)fminf(fmaxf((p.z - lz) * scale, 0.0f), 1023.0f);
      key = cs_spread10(cx) | (cs_spread10(cy) << 1) | (cs_spread10(cz) << 2);
    }
    A[L::phys(i)] = make_uint2(key, (unsigned)gi);
  }
  __syncthreads();

  // ---- phase 2: six stable 5-bit passes over the cluster
  const uint2* chunk = A + t * L::kStride;  // this thread's chunk: IPT consecutive elements of the slice
  for (int pass = 0; pass < kCsPasses; ++pass) {
    const int shift = pass * kCsBits;
    uint2 e[IPT];
    int dg[IPT], rk[IPT];  // digit of each element of the chunk, and how many earlier elements of the chunk share it
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      e[j] = chunk[j];
      dg[j] = (int)((e[j].x >> shift) & (kCsBins - 1));
    }
    // digit counts of the thread's own chunk, from registers: the rank of an element inside the chunk is
    // the number of earlier elements with its digit, and the LAST element of a digit knows the digit's count.
    // Private column of cnt, plain stores, nothing to wait for.
#pragma unroll
    for (int b = 0; b < kCsBins; ++b) cnt[b * kCsThreads + t] = 0;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      int r = 0;
      bool last = true;
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        if (i < j) r += dg[i] == dg[j] ? 1 : 0;
        if (i > j) last = last && dg[i] != dg[j];
      }
      rk[j] = r;
      if (last) cnt[dg[j] * kCsThreads + t] = (unsigned short)(r + 1);
    }
    __syncthreads();
    // exclusive scan of every digit's counts over the threads of the CTA: a warp takes kBinsPerWarp digits,
    // a lane owns kPerLane consecutive threads' counters = kPerLane / 8 128-bit words
    for (int b = warp * kBinsPerWarp; b < (warp + 1) * kBinsPerWarp && b < kCsBins; ++b) {
      static_assert(kPerLane % 8 == 0, "a lane scans whole 128-bit words of 16-bit counters");
      uint4* row = reinterpret_cast<uint4*>(cnt + b * kCsThreads + lane * kPerLane);
      unsigned v[kPerLane / 2];  // two counters per word (little endian)
#pragma unroll
      for (int q = 0; q < kPerLane / 8; ++q) {
        const uint4 w = row[q];
        v[4 * q] = w.x, v[4 * q + 1] = w.y, v[4 * q + 2] = w.z, v[4 * q + 3] = w.w;
      }
      unsigned s = 0;
#pragma unroll
      for (int i = 0; i < kPerLane / 2; ++i) s += (v[i] & 0xffffu) + (v[i] >> 16);
      unsigned inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      unsigned run = inc - s;  // counters of the lanes before this one
#pragma unroll
      for (int i = 0; i < kPerLane / 2; ++i) {
        const unsigned a = v[i] & 0xffffu, c = v[i] >> 16;
        v[i] = run | ((run + a) << 16);
        run += a + c;
      }
#pragma unroll
      for (int q = 0; q < kPerLane / 8; ++q) row[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (lane == 31) tot[b] = inc;
    }
    __syncthreads();
    if (warp == 0) {  // where each digit's run starts in B
      const unsigned lo_v = tot[lane], hi_v = kCsBins > 32 ? tot[(lane + 32) % kCsBins] : 0u;
      unsigned lo_ex, hi_ex;
      cs_scan_bins(lo_v, hi_v, lane, lo_ex, hi_ex);
      lbase[lane] = lo_ex;
      if (kCsBins > 32) lbase[(lane + 32) % kCsBins] = hi_ex;
    }
    __syncthreads();
    // regroup the slice by digit inside the CTA: position = start of the digit's run + elements of the digit in
    // lower threads + rank inside the chunk, so equal digits keep their order (stable)
#pragma unroll
    for (int j = 0; j < IPT; ++j) B[lbase[dg[j]] + cnt[dg[j] * kCsThreads + t] + rk[j]] = e[j];
    cluster.sync();  // every CTA's B and totals are final, and nobody reads its A any more
    for (int i = t; i < kCsBins * kCsCtas; i += kCsThreads) {  // every CTA's total of every digit
      const int b = i / kCsCtas, c = i % kCsCtas;
      all[i] = cluster.map_shared_rank(tot, c)[b];
    }
    __syncthreads();
    if (warp == 0) {
      unsigned row[2] = {0, 0}, before[2] = {0, 0};
#pragma unroll
      for (int hh = 0; hh < (kCsBins > 32 ? 2 : 1); ++hh)
        for (int c = 0; c < kCsCtas; ++c) {
          const unsigned v = all[(lane + 32 * hh) * kCsCtas + c];
          before[hh] += c < rank ? v : 0u;
          row[hh] += v;
        }
      unsigned lo_ex, hi_ex;
      cs_scan_bins(row[0], row[1], lane, lo_ex, hi_ex);
      base[lane] = lo_ex + before[0];  // elements with a smaller digit anywhere + the same digit in lower CTAs
      if (kCsBins > 32) base[(lane + 32) % kCsBins] = hi_ex + before[1];
    }
    __syncthreads();
    // the exchange: a digit's run in B is contiguous in the global order too, so consecutive threads store
    // consecutive 8-byte elements into (mostly) one destination CTA's A: wide, coalesced DSMEM traffic
#pragma unroll
    for (int i = t; i < L::kCap; i += kCsThreads) {
      const uint2 v = B[i];
      const int d = (int)((v.x >> shift) & (kCsBins - 1));
      const unsigned dest = base[d] + ((unsigned)i - lbase[d]);
      const int dc = (int)(dest / (unsigned)L::kCap);
      const int di = L::phys((int)(dest - (unsigned)dc * (unsigned)L::kCap));
      cluster.map_shared_rank(A, dc)[di] = v;
    }
    cluster.sync();  // all elements of this pass are in place
  }

  // ---- phase 3: out, coalesced (padding sits behind position n)
  for (int i = t; i < L::kCap; i += kCsThreads) {
    const int gi = first + i;
    if (gi < n) {
      const uint2 v = A[L::phys(i)];
      keys_out[gi] = v.x;
      vals_out[gi] = (int)v.y;
    }
  }
}

template <int kCsThreads, int IPT>
static int launch_cs(ddlo_runtime* rt, const float4* pts, int n, unsigned* keys_out, int* vals_out, float* lattice) {
  using L = CsLayout<kCsThreads, IPT>;
  auto kern = k_morton_sort_cluster<kCsThreads, IPT>;
  // per device (function attributes are per device): bit set in `configured` once tried, in `usable` if it worked
  static std::atomic<unsigned long long> configured{0}, usable{0};
  const unsigned long long bit = 1ull << (rt->device & 63);
  if (!(configured.load() & bit)) {
    const bool ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes) == cudaSuccess &&
                    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (!ok) (void)cudaGetLastError();
    if (ok) usable.fetch_or(bit);
    configured.fetch_or(bit);
  }
  if (!(usable.load() & bit)) return DDLO_E_UNSUPPORTED;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCsCtas);
  cfg.blockDim = dim3(kCsThreads);
  cfg.dynamicSmemBytes = L::kBytes;
  cfg.stream = rt->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCsCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, pts, n, keys_out, vals_out, lattice) != cudaSuccess) {
    (void)cudaGetLastError();
    usable.fetch_and(~bit);  // e.g. no GPC can host a cluster of 16: never try again on this device
    return DDLO_E_UNSUPPORTED;
  }
  rt->launches += 1;
  return DDLO_OK;
}

#ifndef DDLO_CS_THREADS
#define DDLO_CS_THREADS 512
#endif

// DDLO_OK: keys_out / vals_out / lattice are (being) written on the runtime's stream.
// DDLO_E_UNSUPPORTED (no error text): the cloud is too large or the cluster launch is not possible here.
int morton_sort_cluster(ddlo_runtime* rt, const float4* pts, int n, unsigned* keys_out, int* vals_out, float* lattice) {
  constexpr int T = DDLO_CS_THREADS;  // threads per CTA for the smaller clouds (1024 was measured: index 93 -> 105 us, the block barriers cost more than the extra warps hide)
  if (n <= kCsCtas * T * 2) return launch_cs<T, 2>(rt, pts, n, keys_out, vals_out, lattice);
  if (n <= kCsCtas * T * 4) return launch_cs<T, 4>(rt, pts, n, keys_out, vals_out, lattice);
  if (n <= kCsCtas * 512 * 8) return launch_cs<512, 8>(rt, pts, n, keys_out, vals_out, lattice);
  if (n <= kCsCtas * 512 * 16) return launch_cs<512, 16>(rt, pts, n, keys_out, vals_out, lattice);
  return DDLO_E_UNSUPPORTED;
}

}  // namespace ddlo
