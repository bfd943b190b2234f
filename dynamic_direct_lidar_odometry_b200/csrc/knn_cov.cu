// k-NN queries and per-point covariances.
//   k_knn             KdTreeFLANN::nearestKSearch for a batch of queries        (nanoflann.hpp:146-156)
//   k_cov_from_knn    the arithmetic of NanoGICP::calculate_covariances         (nano_gicp_impl.hpp:392-437)
// The search runs one query per 8-lane sub-warp (knn.cuh) with the k-entry result set and the
// traversal stack in shared memory; the fp64 covariance + regularisation then runs one thread per
// point on the neighbour rows the search left in L2.
#include "common.cuh"
#include "knn.cuh"
#include "math.cuh"

namespace ddlo {

// threads per block (8 lanes per query).  Measured on the C2 step: 512 threads 0.143 ms, 256 0.137, 128 0.1365, 64 0.136,
// 32 0.138 - the kernel is bound by instruction issue (69 % of the slots busy), not by how its blocks are cut.
#ifndef DDLO_KNN_THREADS
#define DDLO_KNN_THREADS 128
#endif
constexpr int kKnnThreads = DDLO_KNN_THREADS;
#ifndef DDLO_KNN_MIN_BLOCKS
#define DDLO_KNN_MIN_BLOCKS 1
#endif
constexpr int kKnnSubs = kKnnThreads / kSubLanes;  // queries per block

// self_mode = 0: query q is queries[q], output row q
// self_mode = 1: query q is the q-th point of the index in Morton order, output row = its original index
// R > 0: k <= 8R, result set in registers (TopKRegSub<R>); R == 0: any k, result set in shared memory
template <int R>
__global__ void __launch_bounds__(kKnnThreads, DDLO_KNN_MIN_BLOCKS) k_knn(IndexView ix, const float4* __restrict__ queries, int nq, int k, int self_mode,
                                                      int* __restrict__ idx_out, float* __restrict__ d_out) {
  extern __shared__ unsigned long long smem_knn[];
  unsigned long long* stacks = smem_knn;  // [kKnnSubs][kStackDepth]
  const Sub sb = make_sub();
  const int sw = threadIdx.x / kSubLanes;
  const int q = blockIdx.x * kKnnSubs + sw;
  const bool active = q < nq;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) v = self_mode ? __ldg(ix.spts + q) : __ldg(queries + q);
  const size_t row = !active ? 0 : (self_mode ? (size_t)__float_as_int(v.w) : (size_t)q);
  unsigned long long* stack = stacks + (size_t)sw * kStackDepth;
  auto search = [&](auto& rs) {
    if (!self_mode) {
      knn_traverse_sub(ix, active, v.x, v.y, v.z, rs, stack, sb);
      return;
    }
    // The query is a point of the cloud: search the node that holds its leaf, then climb.  At each
    // ancestor only the part not searched yet is visited, and the climb stops as soon as the ball
    // of the current k-th distance fits into the ancestor's cube (at the latest at the root).
    int node = active ? __ldg(ix.node_of_point + __float_as_int(v.w)) : 0;
    int skip = -1;
    bool going = active;
    while (__any_sync(kFull, going)) {
      knn_traverse_sub(ix, going, v.x, v.y, v.z, rs, stack, sb, node, skip);
      if (going) {
        const int4 m = __ldg(ix.meta + node);
        if (ball_in_cell(make_ball(ix, v.x, v.y, v.z, rs.worst()), m)) {
          going = false;
        } else {
          skip = node;
          node = m.x;
        }
      }
    }
  };
  if constexpr (R > 0) {
    // Self queries know points of the cloud for free: the consecutive Morton positions around the query.
    // With at least 8R of them the result list starts from that window, ranked in one pass (prefill);
    // otherwise the largest distance among k of them caps the k-th neighbour distance from above.
    constexpr int kWindow = kSubLanes * R;
    const bool window_fill = self_mode && ix.n >= kWindow && k <= kWindow;
    float cap = FLT_MAX;
    if (self_mode && !window_fill && ix.n >= k) {
      float far = 0.0f;
      if (active) {
        const int w0 = min(max(q - k / 2, 0), ix.n - k);
        for (int j = sb.sl; j < k; j += kSubLanes) {
          const float4 p = __ldg(ix.spts + w0 + j);
          far = fmaxf(far, sqdist3_rn(v.x, v.y, v.z, p.x, p.y, p.z));
        }
      }
#pragma unroll
      for (int o = 1; o < kSubLanes; o <<= 1) far = fmaxf(far, __shfl_xor_sync(kFull, far, o));
      if (active && far < FLT_MAX) cap = far;
    }
    TopKRegSub<R> rs;
    rs.init(k, cap);
    if (window_fill) rs.prefill(active, ix.spts, min(max(q - kWindow / 2, 0), ix.n - kWindow), v.x, v.y, v.z, stack, sb);
    search(rs);
    if (active) rs.write_sorted(idx_out + row * k, d_out ? d_out + row * k : nullptr, sb);
  } else {
    float* sd = reinterpret_cast<float*>(smem_knn + (size_t)kKnnSubs * kStackDepth);  // [kKnnSubs][k]
    int* si = reinterpret_cast<int*>(sd + (size_t)kKnnSubs * k);                      // [kKnnSubs][k]
    TopKSub rs;
    rs.init(sd + (size_t)sw * k, si + (size_t)sw * k, k, sb);
    search(rs);
    if (active) rs.write_sorted(idx_out + row * k, d_out ? d_out + row * k : nullptr, sb);
  }
}

// neighbors.colwise() -= neighbors.rowwise().mean(); cov = N N^T / k; regularise   (:392-437), all fp64
__global__ void __launch_bounds__(256) k_cov_from_knn(const float4* __restrict__ pts, int n, const int* __restrict__ knn_idx, int k,
                                                      int method, double* __restrict__ covs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* row = knn_idx + (size_t)i * k;
  double mx = 0.0, my = 0.0, mz = 0.0;
  for (int j = 0; j < k; ++j) {
    DDLO_CHECK_INDEX(__ldg(row + j), n, "k_cov_from_knn: neighbour index");
    const float4 p = __ldg(pts + __ldg(row + j));
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
    const float4 p = __ldg(pts + __ldg(row + j));
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
  double2* out = reinterpret_cast<double2*>(covs + (size_t)i * kCovStride);
  out[0] = make_double2(r.xx, r.xy);
  out[1] = make_double2(r.xz, r.yy);
  out[2] = make_double2(r.yz, r.zz);
}

static int launch_knn(ddlo_runtime* rt, const IndexView& view, const float4* d_queries, int nq, int k, int self_mode, int* d_idx,
                      float* d_d2) {
  const int blocks = (nq + kKnnSubs - 1) / kKnnSubs;
  const size_t stack_bytes = (size_t)kKnnSubs * kStackDepth * 8;
  if (k <= 8) {
    k_knn<1><<<blocks, kKnnThreads, stack_bytes, rt->stream>>>(view, d_queries, nq, k, self_mode, d_idx, d_d2);
  } else if (k <= 16) {
    k_knn<2><<<blocks, kKnnThreads, stack_bytes, rt->stream>>>(view, d_queries, nq, k, self_mode, d_idx, d_d2);
  } else if (k <= 24) {
    k_knn<3><<<blocks, kKnnThreads, stack_bytes, rt->stream>>>(view, d_queries, nq, k, self_mode, d_idx, d_d2);
  } else if (k <= 32) {
    k_knn<4><<<blocks, kKnnThreads, stack_bytes, rt->stream>>>(view, d_queries, nq, k, self_mode, d_idx, d_d2);
  } else {
    const size_t bytes = stack_bytes + (size_t)kKnnSubs * k * 8;
    if (bytes > 200 * 1024) return fail(DDLO_E_UNSUPPORTED, "k too large for the shared-memory result set");
    if (bytes > 48 * 1024) DDLO_CUDA(cudaFuncSetAttribute(k_knn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    k_knn<0><<<blocks, kKnnThreads, bytes, rt->stream>>>(view, d_queries, nq, k, self_mode, d_idx, d_d2);
  }
  rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

int launch_knn_queries(ddlo_cloud* c, const float4* d_queries, int nq, int k, int* d_idx, float* d_d2) {
  return launch_knn(c->rt, c->view, d_queries, nq, k, 0, d_idx, d_d2);
}

int launch_covariances(ddlo_cloud* c, int k, int method, double* d_covs) {
  ddlo_runtime* rt = c->rt;
  int* d_idx = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_idx), (size_t)c->n * k * sizeof(int), rt->stream));
  int rc = launch_knn(rt, c->view, nullptr, c->n, k, 1, d_idx, nullptr);
  if (rc == DDLO_OK) {
    k_cov_from_knn<<<(c->n + 255) / 256, 256, 0, rt->stream>>>(c->pts, c->n, d_idx, k, method, d_covs);
    rt->launches += 1;
  }
  cudaFreeAsync(d_idx, rt->stream);
  if (rc != DDLO_OK) return rc;
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

}  // namespace ddlo
