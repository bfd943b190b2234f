// k-NN queries and per-point covariances.
//   k_knn_queries   KdTreeFLANN::nearestKSearch for a batch of external queries  (nanoflann.hpp:146-156)
//   k_covariances   NanoGICP::calculate_covariances                              (nano_gicp_impl.hpp:374-441)
// One thread per query; the k-entry result set of each thread lives in shared memory (entry-major,
// bank-conflict free); the fp64 covariance and its regularisation are fused behind the search so
// neighbour indices never travel through HBM.
#include "common.cuh"
#include "knn.cuh"
#include "math.cuh"

namespace ddlo {

constexpr int kKnnThreads = 128;

__global__ void __launch_bounds__(kKnnThreads) k_knn_queries(IndexView ix, const float4* __restrict__ queries, int nq, int k,
                                                              int* __restrict__ idx_out, float* __restrict__ d_out) {
  extern __shared__ unsigned char smem[];
  float* sd = reinterpret_cast<float*>(smem);
  int* si = reinterpret_cast<int*>(smem + sizeof(float) * k * kKnnThreads);
  const int q = blockIdx.x * kKnnThreads + threadIdx.x;
  TopKShared rs;
  rs.init(sd, si, k, kKnnThreads, threadIdx.x);
  if (q >= nq) return;
  const float4 v = queries[q];
  knn_traverse(ix, v.x, v.y, v.z, rs);
  for (int j = 0; j < k; ++j) {
    const int id = rs.idx[j * rs.stride];
    idx_out[(size_t)q * k + j] = id;
    d_out[(size_t)q * k + j] = id < 0 ? __int_as_float(0x7f800000) : rs.d[j * rs.stride];
  }
}

// thread t handles the t-th point in MORTON order (neighbouring threads walk neighbouring paths)
__global__ void __launch_bounds__(kKnnThreads) k_covariances(IndexView ix, const float4* __restrict__ pts, int k, int method,
                                                              double* __restrict__ covs) {
  extern __shared__ unsigned char smem[];
  float* sd = reinterpret_cast<float*>(smem);
  int* si = reinterpret_cast<int*>(smem + sizeof(float) * k * kKnnThreads);
  const int s = blockIdx.x * kKnnThreads + threadIdx.x;
  TopKShared rs;
  rs.init(sd, si, k, kKnnThreads, threadIdx.x);
  if (s >= ix.n) return;
  const float4 v = __ldg(ix.spts + s);
  const int self = __float_as_int(v.w);
  knn_traverse(ix, v.x, v.y, v.z, rs);

  // neighbors.colwise() -= neighbors.rowwise().mean(); cov = N N^T / k   (:392-399), all fp64
  double mx = 0.0, my = 0.0, mz = 0.0;
  for (int j = 0; j < k; ++j) {
    const float4 p = __ldg(pts + rs.idx[j * rs.stride]);
    mx += (double)p.x;
    my += (double)p.y;
    mz += (double)p.z;
  }
  const double kd = (double)k;
  mx /= kd;
  my /= kd;
  mz /= kd;
  Sym3 c = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < k; ++j) {
    const float4 p = __ldg(pts + rs.idx[j * rs.stride]);
    const double dx = (double)p.x - mx, dy = (double)p.y - my, dz = (double)p.z - mz;
    c.xx += dx * dx;
    c.xy += dx * dy;
    c.xz += dx * dz;
    c.yy += dy * dy;
    c.yz += dy * dz;
    c.zz += dz * dz;
  }
  c.xx /= kd, c.xy /= kd, c.xz /= kd, c.yy /= kd, c.yz /= kd, c.zz /= kd;
  const Sym3 r = regularize_cov(c, method);
  double2* out = reinterpret_cast<double2*>(covs + (size_t)self * kCovStride);
  out[0] = make_double2(r.xx, r.xy);
  out[1] = make_double2(r.xz, r.yy);
  out[2] = make_double2(r.yz, r.zz);
}

int launch_knn_queries(ddlo_cloud* c, const float4* d_queries, int nq, int k, int* d_idx, float* d_d2) {
  const size_t smem = (size_t)k * kKnnThreads * 8;
  if (smem > 200 * 1024) return fail(DDLO_E_UNSUPPORTED, "k too large for the shared-memory result set");
  if (smem > 48 * 1024) DDLO_CUDA(cudaFuncSetAttribute(k_knn_queries, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_knn_queries<<<(nq + kKnnThreads - 1) / kKnnThreads, kKnnThreads, smem, c->rt->stream>>>(c->view, d_queries, nq, k, d_idx, d_d2);
  c->rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

int launch_covariances(ddlo_cloud* c, int k, int method, double* d_covs) {
  const size_t smem = (size_t)k * kKnnThreads * 8;
  if (smem > 200 * 1024) return fail(DDLO_E_UNSUPPORTED, "k too large for the shared-memory result set");
  if (smem > 48 * 1024) DDLO_CUDA(cudaFuncSetAttribute(k_covariances, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_covariances<<<(c->n + kKnnThreads - 1) / kKnnThreads, kKnnThreads, smem, c->rt->stream>>>(c->view, c->pts, k, method, d_covs);
  c->rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

}  // namespace ddlo
