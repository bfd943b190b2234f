// extern "C" surface of libddlo_gicp_b200.so (declared in include/ddlo_gicp.h).
// Host-side bookkeeping only: handles, reference counts, uploads/downloads and the state rules of
// NanoGICP (nano_gicp_impl.hpp:98-196).  All arithmetic happens in the kernels of index.cu,
// knn_cov.cu and gicp.cu; there is no CPU implementation of any of it in this library.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "common.cuh"
#include "gicp.cuh"
#include "engine.cuh"

namespace ddlo {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int build_index(ddlo_cloud* c);
int launch_knn_queries(ddlo_cloud* c, const float4* d_queries, int nq, int k, int* d_idx, float* d_d2);
int launch_covariances(ddlo_cloud* c, int k, int method, double* d_covs);
int index_nonfinite_count(ddlo_cloud* c, int* count);
int voxel_filter_device(ddlo_runtime* rt, const float4* pts, int n, const float leaf[3], float4** d_out, int* n_out);
int crop_box_device(ddlo_runtime* rt, const float4* pts, int n, const float lo[3], const float hi[3], int negative, int keep_organized,
                    float4** d_out, int* n_out);
int extract_stride_device(ddlo_runtime* rt, const float4* pts, int n, int width, int height, int row_stride, int col_stride, float4** d_out);
int residual_image_device(ddlo_runtime* rt, const float4* pts, const float* sqd, int n, int w, int h, double a_min, double a_max, float4* d_out);
int segment_scan_device(ddlo_runtime* rt, const ddlo_segmentation_params& prm, const float* T16, const float* d_scan, int stride_floats,
                        const float* d_residuals, int res_stride, int* d_label, float* d_range, signed char* d_ground, double* d_avg_by_label,
                        int* d_label_count);

// raw strided host points -> float4 (x, y, z, 1)
__global__ void __launch_bounds__(256) k_repack(const unsigned char* __restrict__ raw, int n, int stride, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
  out[i] = make_float4(p[0], p[1], p[2], 1.0f);
}

// Eigen::Matrix4d per point <-> 6 doubles per point
__global__ void __launch_bounds__(256) k_cov_pack(const double* __restrict__ m16, int n, double* __restrict__ c6) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* m = m16 + (size_t)i * 16;  // column-major: (r,c) at m[4c + r]
  double* o = c6 + (size_t)i * 6;
  o[0] = m[0];
  o[1] = 0.5 * (m[4] + m[1]);
  o[2] = 0.5 * (m[8] + m[2]);
  o[3] = m[5];
  o[4] = 0.5 * (m[9] + m[6]);
  o[5] = m[10];
}
__global__ void __launch_bounds__(256) k_cov_unpack(const double* __restrict__ c6, int n, double* __restrict__ m16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* s = c6 + (size_t)i * 6;
  double* m = m16 + (size_t)i * 16;
  m[0] = s[0], m[1] = s[1], m[2] = s[2], m[3] = 0.0;
  m[4] = s[1], m[5] = s[3], m[6] = s[4], m[7] = 0.0;
  m[8] = s[2], m[9] = s[4], m[10] = s[5], m[11] = 0.0;
  m[12] = 0.0, m[13] = 0.0, m[14] = 0.0, m[15] = 0.0;
}
__global__ void __launch_bounds__(256) k_sqrt_f2d(const float* __restrict__ in, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = sqrt((double)in[i]);
}
__global__ void __launch_bounds__(256) k_fill(float4* p, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Concatenation of up to kConcatParts device arrays of 16-byte words in ONE launch (a submap is 10 - 30 keyframes;
// one cudaMemcpyAsync per part costs more in launches than the copy itself)
constexpr int kConcatParts = 32;
struct ConcatParts {
  const uint4* src[kConcatParts];
  unsigned long long begin[kConcatParts + 1];  // in 16-byte words
  int m;
};
__global__ void __launch_bounds__(256) k_concat(ConcatParts p, uint4* __restrict__ dst) {
  const unsigned long long total = p.begin[p.m];
  for (unsigned long long w = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; w < total; w += (unsigned long long)gridDim.x * blockDim.x) {
    int part = 0;
#pragma unroll 1
    for (int lo = 0, hi = p.m; lo < hi;) {  // the part that holds word w
      const int mid = (lo + hi) >> 1;
      if (p.begin[mid + 1] <= w)
        lo = mid + 1;
      else
        hi = mid;
      part = lo;
    }
    dst[w] = p.src[part][w - p.begin[part]];
  }
}

// parts[i]: n_words[i] 16-byte words at src[i]; empty parts are skipped
static int concat_words(ddlo_runtime* rt, const void* const* src, const size_t* n_words, int m, void* dst) {
  size_t off = 0;
  int i = 0;
  while (i < m) {
    ConcatParts p;
    p.m = 0;
    p.begin[0] = 0;
    unsigned long long words = 0;
    for (; i < m && p.m < kConcatParts; ++i) {
      if (n_words[i] == 0) continue;
      p.src[p.m] = static_cast<const uint4*>(src[i]);
      words += n_words[i];
      p.begin[++p.m] = words;
    }
    if (p.m == 0) break;
    const int blocks = (int)std::min<unsigned long long>((words + 255) / 256, (unsigned long long)rt->num_sms * 16);
    k_concat<<<blocks, 256, 0, rt->stream>>>(p, static_cast<uint4*>(dst) + off);
    rt->launches += 1;
    DDLO_CUDA(cudaGetLastError());
    off += words;
  }
  return DDLO_OK;
}

int use_device(const ddlo_runtime* rt) {
  DDLO_CUDA(cudaSetDevice(rt->device));
  return DDLO_OK;
}

int ensure_pinned(ddlo_runtime* rt, size_t bytes) {
  if (rt->h_pinned_bytes >= bytes) return DDLO_OK;
  if (rt->h_pinned) {
    DDLO_CUDA(cudaStreamSynchronize(rt->stream));
    DDLO_CUDA(cudaFreeHost(rt->h_pinned));
    rt->h_pinned = nullptr;
    rt->h_pinned_bytes = 0;
  }
  const size_t want = std::max<size_t>(bytes, 1u << 16);
  DDLO_CUDA(cudaMallocHost(&rt->h_pinned, want));
  rt->h_pinned_bytes = want;
  return DDLO_OK;
}

static void runtime_teardown(ddlo_runtime* rt) {
  cudaSetDevice(rt->device);
  if (rt->stream) cudaStreamSynchronize(rt->stream);
  if (rt->d_scratch) cudaFree(rt->d_scratch);
  if (rt->flush_buf) cudaFree(rt->flush_buf);
  if (rt->h_pinned) cudaFreeHost(rt->h_pinned);
  if (rt->ev0) cudaEventDestroy(rt->ev0);
  if (rt->ev1) cudaEventDestroy(rt->ev1);
  if (rt->ev_fork) cudaEventDestroy(rt->ev_fork);
  if (rt->ev_join) cudaEventDestroy(rt->ev_join);
  for (auto& e : rt->slots)
    if (e) cudaEventDestroy(e);
  if (rt->ev_fork2) cudaEventDestroy(rt->ev_fork2);
  if (rt->ev_join2) cudaEventDestroy(rt->ev_join2);
  for (cudaStream_t q : {rt->side, rt->side2})
    if (q) {
      cudaStreamSynchronize(q);
      cudaStreamDestroy(q);
    }
  if (rt->stream) cudaStreamDestroy(rt->stream);
  delete rt;
}
void runtime_retain(ddlo_runtime* rt) { rt->refs.fetch_add(1); }
void runtime_release(ddlo_runtime* rt) {
  if (rt->refs.fetch_sub(1) == 1) runtime_teardown(rt);
}

static void cloud_free(ddlo_cloud* c) {
  cudaSetDevice(c->rt->device);
  cudaStream_t st = c->rt->stream;
  if (c->pts) cudaFreeAsync(c->pts, st);
  if (c->spts) cudaFreeAsync(c->spts, st);
  if (c->nodes) cudaFreeAsync(c->nodes, st);
  if (c->meta) cudaFreeAsync(c->meta, st);
  if (c->node_of_point) cudaFreeAsync(c->node_of_point, st);
  if (c->lattice) cudaFreeAsync(c->lattice, st);
  runtime_release(c->rt);
  delete c;
}
static void covs_free(ddlo_covs* v) {
  cudaSetDevice(v->rt->device);
  if (v->c) cudaFreeAsync(v->c, v->rt->stream);
  if (v->sorted) cudaFreeAsync(v->sorted, v->rt->stream);
  if (v->sorted_for && v->sorted_for->refs.fetch_sub(1) == 1) cloud_free(v->sorted_for);
  runtime_release(v->rt);
  delete v;
}

void cloud_set(ddlo_cloud*& slot, ddlo_cloud* c) {
  if (c) c->refs.fetch_add(1);
  if (slot && slot->refs.fetch_sub(1) == 1) cloud_free(slot);
  slot = c;
}
void covs_set(ddlo_covs*& slot, ddlo_covs* v) {
  if (v) v->refs.fetch_add(1);
  if (slot && slot->refs.fetch_sub(1) == 1) covs_free(slot);
  slot = v;
}

}  // namespace ddlo

using namespace ddlo;

static inline void set_cloud(ddlo_cloud*& slot, ddlo_cloud* c) { cloud_set(slot, c); }
static inline void set_covs(ddlo_covs*& slot, ddlo_covs* v) { covs_set(slot, v); }

namespace ddlo {

int cloud_new(ddlo_runtime* rt, int n, ddlo_cloud** out) {
  ddlo_cloud* c = new (std::nothrow) ddlo_cloud();
  if (!c) return fail(DDLO_E_INVALID, "out of host memory");
  c->rt = rt;
  c->n = n;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&c->pts), std::max<size_t>(1, (size_t)n) * sizeof(float4), rt->stream);
  if (e != cudaSuccess) {
    delete c;
    return fail(DDLO_E_CUDA, std::string("cudaMallocAsync(cloud): ") + cudaGetErrorString(e));
  }
  runtime_retain(rt);
  *out = c;
  return DDLO_OK;
}

// a cloud handle around a device buffer that is already filled (takes ownership)
int cloud_adopt(ddlo_runtime* rt, float4* pts, int n, ddlo_cloud** out) {
  ddlo_cloud* c = new (std::nothrow) ddlo_cloud();
  if (!c) {
    cudaFreeAsync(pts, rt->stream);
    return fail(DDLO_E_INVALID, "out of host memory");
  }
  c->rt = rt;
  c->n = n;
  c->pts = pts;
  runtime_retain(rt);
  *out = c;
  return DDLO_OK;
}

int covs_new(ddlo_runtime* rt, int n, ddlo_covs** out) {
  ddlo_covs* v = new (std::nothrow) ddlo_covs();
  if (!v) return fail(DDLO_E_INVALID, "out of host memory");
  v->rt = rt;
  v->n = n;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&v->c), std::max<size_t>(1, (size_t)n) * kCovStride * sizeof(double), rt->stream);
  if (e != cudaSuccess) {
    delete v;
    return fail(DDLO_E_CUDA, std::string("cudaMallocAsync(covs): ") + cudaGetErrorString(e));
  }
  runtime_retain(rt);
  *out = v;
  return DDLO_OK;
}

}  // namespace ddlo

extern "C" {

int ddlo_abi_version(void) { return DDLO_ABI_VERSION; }
const char* ddlo_last_error(void) { return g_err.c_str(); }

int ddlo_device_count(int* count) {
  if (!count) return fail(DDLO_E_INVALID, "count is null");
  DDLO_CUDA(cudaGetDeviceCount(count));
  return DDLO_OK;
}

// ---- runtime --------------------------------------------------------------------------------------
static int runtime_init(ddlo_runtime* rt, int device);
int ddlo_runtime_create(int device, ddlo_runtime** out) {
  if (!out) return fail(DDLO_E_INVALID, "out is null");
  *out = nullptr;
  int count = 0;
  DDLO_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(DDLO_E_CUDA, "no such CUDA device (this library has no CPU fallback)");
  DDLO_CUDA(cudaSetDevice(device));
  ddlo_runtime* rt = new (std::nothrow) ddlo_runtime();
  if (!rt) return fail(DDLO_E_INVALID, "out of host memory");
  rt->device = device;
  const int rc = runtime_init(rt, device);
  if (rc != DDLO_OK) {
    const std::string why = g_err;
    ddlo_runtime_destroy(rt);  // releases whatever was created
    return fail(rc, why);
  }
  *out = rt;
  return DDLO_OK;
}

static int runtime_init(ddlo_runtime* rt, int device) {
  DDLO_CUDA(cudaStreamCreateWithFlags(&rt->stream, cudaStreamNonBlocking));
  DDLO_CUDA(cudaStreamCreateWithFlags(&rt->side, cudaStreamNonBlocking));
  DDLO_CUDA(cudaStreamCreateWithFlags(&rt->side2, cudaStreamNonBlocking));
  DDLO_CUDA(cudaEventCreateWithFlags(&rt->ev_fork2, cudaEventDisableTiming));
  DDLO_CUDA(cudaEventCreateWithFlags(&rt->ev_join2, cudaEventDisableTiming));
  DDLO_CUDA(cudaEventCreateWithFlags(&rt->ev_fork, cudaEventDisableTiming));
  DDLO_CUDA(cudaEventCreateWithFlags(&rt->ev_join, cudaEventDisableTiming));
  DDLO_CUDA(cudaEventCreate(&rt->ev0));
  DDLO_CUDA(cudaEventCreate(&rt->ev1));
  for (auto& e : rt->slots) DDLO_CUDA(cudaEventCreate(&e));
  DDLO_CUDA(cudaDeviceGetAttribute(&rt->num_sms, cudaDevAttrMultiProcessorCount, device));
  cudaMemPool_t pool;
  DDLO_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t keep = ~0ull;  // keep freed blocks cached: handle churn must not hit cudaMalloc
  DDLO_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  // Grow the pool once, now, instead of in the middle of a frame: the first allocation of a size the pool has
  // never held maps new physical memory (tens of milliseconds for the ~100 MB node array of a 650k-point submap).
  // DDLO_POOL_PREWARM_MB overrides the default of 1 GiB (0 disables).
  {
    size_t mb = 1024;
    if (const char* e = std::getenv("DDLO_POOL_PREWARM_MB")) mb = (size_t)std::max(0L, std::atol(e));
    if (mb > 0) {
      void* warm = nullptr;
      if (cudaMallocAsync(&warm, mb << 20, rt->stream) == cudaSuccess)
        cudaFreeAsync(warm, rt->stream);
      else
        (void)cudaGetLastError();  // not enough memory for the warm-up: carry on without it
    }
  }
  DDLO_CUDA(cudaMalloc(&rt->d_scratch, 1u << 20));
  rt->d_scratch_bytes = 1u << 20;
  DDLO_CUDA(cudaMemsetAsync(rt->d_scratch, 0, rt->d_scratch_bytes, rt->stream));  // the index build keeps its ticket word here
  DDLO_TRY(ensure_pinned(rt, 1u << 16));
  int per_sm = 0;
  DDLO_TRY(gicp_max_coop_blocks(device, &per_sm));
  rt->max_coop_blocks_align = per_sm * rt->num_sms;
  if (rt->max_coop_blocks_align <= 0) return fail(DDLO_E_CUDA, "align kernel does not fit on this device");
  rt->align_blocks_limit = rt->max_coop_blocks_align;
  return DDLO_OK;
}

int ddlo_runtime_destroy(ddlo_runtime* rt) {
  if (!rt) return DDLO_OK;
  cudaSetDevice(rt->device);
  if (rt->stream) cudaStreamSynchronize(rt->stream);
  runtime_release(rt);  // the stream and the scratch buffers go with the last handle made from this runtime
  return DDLO_OK;
}

int ddlo_runtime_set_align_blocks(ddlo_runtime* rt, int max_blocks) {
  if (!rt) return fail(DDLO_E_INVALID, "runtime is null");
  rt->align_blocks_limit = max_blocks <= 0 ? rt->max_coop_blocks_align : std::min(max_blocks, rt->max_coop_blocks_align);
  return DDLO_OK;
}

int ddlo_runtime_synchronize(ddlo_runtime* rt) {
  if (!rt) return fail(DDLO_E_INVALID, "runtime is null");
  DDLO_TRY(use_device(rt));
  DDLO_CUDA(cudaStreamSynchronize(rt->stream));
  return DDLO_OK;
}

int ddlo_runtime_timer_begin(ddlo_runtime* rt) {
  if (!rt) return fail(DDLO_E_INVALID, "runtime is null");
  DDLO_TRY(use_device(rt));
  DDLO_CUDA(cudaEventRecord(rt->ev0, rt->stream));
  return DDLO_OK;
}

int ddlo_runtime_timer_end(ddlo_runtime* rt, float* elapsed_ms) {
  if (!rt || !elapsed_ms) return fail(DDLO_E_INVALID, "null argument");
  DDLO_TRY(use_device(rt));
  DDLO_CUDA(cudaEventRecord(rt->ev1, rt->stream));
  DDLO_CUDA(cudaEventSynchronize(rt->ev1));
  DDLO_CUDA(cudaEventElapsedTime(elapsed_ms, rt->ev0, rt->ev1));
  return DDLO_OK;
}

int ddlo_runtime_launch_count(ddlo_runtime* rt, long long* count) {
  if (!rt || !count) return fail(DDLO_E_INVALID, "null argument");
  *count = rt->launches;
  return DDLO_OK;
}

int ddlo_runtime_flush_l2(ddlo_runtime* rt, size_t bytes) {
  if (!rt) return fail(DDLO_E_INVALID, "runtime is null");
  DDLO_TRY(use_device(rt));
  if (rt->flush_bytes < bytes) {
    if (rt->flush_buf) {
      DDLO_CUDA(cudaStreamSynchronize(rt->stream));
      DDLO_CUDA(cudaFree(rt->flush_buf));
      rt->flush_buf = nullptr;
      rt->flush_bytes = 0;
    }
    DDLO_CUDA(cudaMalloc(&rt->flush_buf, bytes));
    rt->flush_bytes = bytes;
  }
  k_fill<<<rt->num_sms * 8, 256, 0, rt->stream>>>(static_cast<float4*>(rt->flush_buf), bytes / 16);
  DDLO_CUDA(cudaGetLastError());
  return DDLO_OK;
}

int ddlo_runtime_event_record(ddlo_runtime* rt, int slot) {
  if (!rt || slot < 0 || slot >= 16) return fail(DDLO_E_INVALID, "bad runtime or event slot");
  DDLO_TRY(use_device(rt));
  DDLO_CUDA(cudaEventRecord(rt->slots[slot], rt->stream));
  return DDLO_OK;
}

int ddlo_runtime_event_elapsed(ddlo_runtime* rt, int a, int b, float* elapsed_ms) {
  if (!rt || !elapsed_ms || a < 0 || a >= 16 || b < 0 || b >= 16) return fail(DDLO_E_INVALID, "bad runtime or event slot");
  DDLO_TRY(use_device(rt));
  DDLO_CUDA(cudaEventSynchronize(rt->slots[b]));
  DDLO_CUDA(cudaEventElapsedTime(elapsed_ms, rt->slots[a], rt->slots[b]));
  return DDLO_OK;
}

int ddlo_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(DDLO_E_INVALID, "out is null");
  DDLO_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return DDLO_OK;
}
int ddlo_host_free(void* p) {
  if (p) DDLO_CUDA(cudaFreeHost(p));
  return DDLO_OK;
}


// ---- clouds ---------------------------------------------------------------------------------------
int ddlo_cloud_create(ddlo_runtime* rt, const float* xyz, int n, int stride_bytes, ddlo_cloud** out) {
  if (!rt || !out || (n > 0 && !xyz)) return fail(DDLO_E_INVALID, "null argument");
  if (n < 0 || stride_bytes < 12 || (stride_bytes & 3)) return fail(DDLO_E_INVALID, "bad n or stride");
  *out = nullptr;
  DDLO_TRY(use_device(rt));
  ddlo_cloud* c = nullptr;
  DDLO_TRY(cloud_new(rt, n, &c));
  if (n > 0) {
    const size_t raw_bytes = (size_t)(n - 1) * stride_bytes + 12;
    unsigned char* d_raw = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&d_raw), raw_bytes, rt->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_raw, xyz, raw_bytes, cudaMemcpyHostToDevice, rt->stream);
    if (e == cudaSuccess) {
      k_repack<<<(n + 255) / 256, 256, 0, rt->stream>>>(d_raw, n, stride_bytes, c->pts);
      rt->launches += 1;
      e = cudaGetLastError();
    }
    if (d_raw) cudaFreeAsync(d_raw, rt->stream);
    if (e != cudaSuccess) {
      cloud_free(c);
      return fail(DDLO_E_CUDA, std::string("cloud upload: ") + cudaGetErrorString(e));
    }
  }
  *out = c;
  return DDLO_OK;
}

int ddlo_cloud_create_from_device(ddlo_runtime* rt, const void* d_xyzw, int n, ddlo_cloud** out) {
  if (!rt || !out || (n > 0 && !d_xyzw)) return fail(DDLO_E_INVALID, "null argument");
  if (n < 0) return fail(DDLO_E_INVALID, "bad n");
  *out = nullptr;
  DDLO_TRY(use_device(rt));
  ddlo_cloud* c = nullptr;
  DDLO_TRY(cloud_new(rt, n, &c));
  if (n > 0) {
    k_repack<<<(n + 255) / 256, 256, 0, rt->stream>>>(static_cast<const unsigned char*>(d_xyzw), n, 16, c->pts);
    rt->launches += 1;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      cloud_free(c);
      return fail(DDLO_E_CUDA, std::string("cloud copy: ") + cudaGetErrorString(e));
    }
  }
  *out = c;
  return DDLO_OK;
}

int ddlo_cloud_retain(ddlo_cloud* c) {
  if (!c) return fail(DDLO_E_INVALID, "cloud is null");
  c->refs.fetch_add(1);
  return DDLO_OK;
}
int ddlo_cloud_release(ddlo_cloud* c) {
  if (!c) return DDLO_OK;
  if (c->refs.fetch_sub(1) == 1) cloud_free(c);
  return DDLO_OK;
}
int ddlo_cloud_size(const ddlo_cloud* c, int* n) {
  if (!c || !n) return fail(DDLO_E_INVALID, "null argument");
  *n = c->n;
  return DDLO_OK;
}
int ddlo_cloud_download(ddlo_cloud* c, float* xyzw_out) {
  if (!c || !xyzw_out) return fail(DDLO_E_INVALID, "null argument");
  DDLO_TRY(use_device(c->rt));
  DDLO_CUDA(cudaMemcpyAsync(xyzw_out, c->pts, (size_t)c->n * sizeof(float4), cudaMemcpyDeviceToHost, c->rt->stream));
  DDLO_CUDA(cudaStreamSynchronize(c->rt->stream));
  return DDLO_OK;
}
int ddlo_cloud_build_index(ddlo_cloud* c) {
  if (!c) return fail(DDLO_E_INVALID, "cloud is null");
  DDLO_TRY(use_device(c->rt));
  return build_index(c);
}
int ddlo_cloud_has_index(const ddlo_cloud* c, int* has) {
  if (!c || !has) return fail(DDLO_E_INVALID, "null argument");
  *has = c->has_index ? 1 : 0;
  return DDLO_OK;
}

int ddlo_cloud_knn(ddlo_cloud* c, const float* queries, int nq, int qstride_bytes, int k, int* idx, float* sqdist, int* counts) {
  if (!c || (nq > 0 && (!queries || !idx || !sqdist))) return fail(DDLO_E_INVALID, "null argument");
  if (nq < 0 || k < 1 || qstride_bytes < 12 || (qstride_bytes & 3)) return fail(DDLO_E_INVALID, "bad nq, k or stride");
  if (c->n <= 0) return fail(DDLO_E_EMPTY, "kNN on an empty cloud");
  ddlo_runtime* rt = c->rt;
  DDLO_TRY(use_device(rt));
  DDLO_TRY(build_index(c));
  if (nq == 0) return DDLO_OK;
  cudaStream_t st = rt->stream;
  const size_t raw_bytes = (size_t)(nq - 1) * qstride_bytes + 12;
  unsigned char* d_raw = nullptr;
  float4* d_q = nullptr;
  int* d_idx = nullptr;
  float* d_d2 = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_raw), raw_bytes, st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_q), (size_t)nq * sizeof(float4), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_idx), (size_t)nq * k * sizeof(int), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_d2), (size_t)nq * k * sizeof(float), st));
  DDLO_CUDA(cudaMemcpyAsync(d_raw, queries, raw_bytes, cudaMemcpyHostToDevice, st));
  k_repack<<<(nq + 255) / 256, 256, 0, st>>>(d_raw, nq, qstride_bytes, d_q);
  rt->launches += 1;
  int rc = launch_knn_queries(c, d_q, nq, k, d_idx, d_d2);
  if (rc == DDLO_OK) {
    DDLO_CUDA(cudaMemcpyAsync(idx, d_idx, (size_t)nq * k * sizeof(int), cudaMemcpyDeviceToHost, st));
    DDLO_CUDA(cudaMemcpyAsync(sqdist, d_d2, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  cudaFreeAsync(d_raw, st);
  cudaFreeAsync(d_q, st);
  cudaFreeAsync(d_idx, st);
  cudaFreeAsync(d_d2, st);
  DDLO_CUDA(cudaStreamSynchronize(st));
  if (rc != DDLO_OK) return rc;
  int bad = 0;
  DDLO_TRY(index_nonfinite_count(c, &bad));
  if (bad) return fail(DDLO_E_NONFINITE, "kNN: the cloud holds NaN or Inf coordinates");
  if (counts)
    for (int i = 0; i < nq; ++i) counts[i] = std::min(k, c->n);
  return DDLO_OK;
}

int ddlo_cloud_transform(ddlo_cloud* c, const float* T16, ddlo_cloud** out) {
  if (!c || !T16 || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = nullptr;
  DDLO_TRY(use_device(c->rt));
  ddlo_cloud* r = nullptr;
  DDLO_TRY(cloud_new(c->rt, c->n, &r));
  if (c->n > 0) {
    int rc = launch_transform_cloud(c->rt, c->pts, c->n, T16, r->pts);
    if (rc != DDLO_OK) {
      cloud_free(r);
      return rc;
    }
  }
  *out = r;
  return DDLO_OK;
}


int ddlo_cloud_voxel_filter(ddlo_cloud* c, float leaf_x, float leaf_y, float leaf_z, ddlo_cloud** out) {
  if (!c || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = nullptr;
  DDLO_TRY(use_device(c->rt));
  const float leaf[3] = {leaf_x, leaf_y, leaf_z};
  float4* d = nullptr;
  int n = 0;
  DDLO_TRY(voxel_filter_device(c->rt, c->pts, c->n, leaf, &d, &n));
  return cloud_adopt(c->rt, d, n, out);
}

int ddlo_cloud_crop_box(ddlo_cloud* c, const float* min_xyz, const float* max_xyz, int negative, int keep_organized, ddlo_cloud** out) {
  if (!c || !min_xyz || !max_xyz || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = nullptr;
  DDLO_TRY(use_device(c->rt));
  float4* d = nullptr;
  int n = 0;
  DDLO_TRY(crop_box_device(c->rt, c->pts, c->n, min_xyz, max_xyz, negative, keep_organized, &d, &n));
  return cloud_adopt(c->rt, d, n, out);
}

int ddlo_cloud_extract_stride(ddlo_cloud* c, int width, int height, int row_stride, int col_stride, ddlo_cloud** out) {
  if (!c || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = nullptr;
  if (width < 1 || height < 1 || row_stride < 1 || col_stride < 1) return fail(DDLO_E_INVALID, "extract_stride: sizes and strides must be positive");
  if ((long long)width * height > c->n) return fail(DDLO_E_SIZE, "extract_stride: the cloud is smaller than width x height (not an organised scan of that shape)");
  DDLO_TRY(use_device(c->rt));
  float4* d = nullptr;
  DDLO_TRY(extract_stride_device(c->rt, c->pts, c->n, width, height, row_stride, col_stride, &d));
  return cloud_adopt(c->rt, d, c->n, out);
}

int ddlo_cloud_concat(ddlo_runtime* rt, ddlo_cloud* const* parts, int m, ddlo_cloud** out) {
  if (!rt || !out || (m > 0 && !parts)) return fail(DDLO_E_INVALID, "null argument");
  *out = nullptr;
  DDLO_TRY(use_device(rt));
  long long total = 0;
  for (int i = 0; i < m; ++i) {
    if (!parts[i] || parts[i]->rt != rt) return fail(DDLO_E_INVALID, "concat: null part or part of another runtime");
    total += parts[i]->n;
  }
  if (total > std::numeric_limits<int>::max()) return fail(DDLO_E_UNSUPPORTED, "concat: too many points");
  ddlo_cloud* r = nullptr;
  DDLO_TRY(cloud_new(rt, (int)total, &r));
  std::vector<const void*> src(m);
  std::vector<size_t> words(m);
  for (int i = 0; i < m; ++i) src[i] = parts[i]->pts, words[i] = (size_t)parts[i]->n;  // a point is one 16-byte word
  const int rc = concat_words(rt, src.data(), words.data(), m, r->pts);
  if (rc != DDLO_OK) {
    cloud_free(r);
    return rc;
  }
  *out = r;
  return DDLO_OK;
}


// Make a finished cloud (points uploaded, index built) readable from other runtimes (streams) of the same device:
// synchronises the owner's stream once; from then on the handle is immutable.
int ddlo_cloud_share(ddlo_cloud* c) {
  if (!c) return fail(DDLO_E_INVALID, "cloud is null");
  DDLO_TRY(use_device(c->rt));
  if (c->n > 0) DDLO_TRY(build_index(c));
  DDLO_CUDA(cudaStreamSynchronize(c->rt->stream));
  c->shared = true;
  return DDLO_OK;
}

// ---- covariances ------------------------------------------------------------------------------------
int ddlo_covs_compute(ddlo_cloud* c, int k, int regularization_method, ddlo_covs** out) {
  if (!c || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = nullptr;
  if (c->n <= 0) return fail(DDLO_E_EMPTY, "covariances of an empty cloud");
  if (k < 1) return fail(DDLO_E_INVALID, "k must be >= 1");
  if (c->n < k) return fail(DDLO_E_TOO_FEW, "cloud has fewer points than k_correspondences");
  if (regularization_method < DDLO_REG_NONE || regularization_method > DDLO_REG_FROBENIUS)
    return fail(DDLO_E_INVALID, "unknown regularization method");
  DDLO_TRY(use_device(c->rt));
  DDLO_TRY(build_index(c));
  ddlo_covs* v = nullptr;
  DDLO_TRY(covs_new(c->rt, c->n, &v));
  int rc = launch_covariances(c, k, regularization_method, v->c);
  if (rc != DDLO_OK) {
    covs_free(v);
    return rc;
  }
  *out = v;
  return DDLO_OK;
}

int ddlo_covs_from_host(ddlo_runtime* rt, const double* mat4x4, int n, ddlo_covs** out) {
  if (!rt || !out || (n > 0 && !mat4x4)) return fail(DDLO_E_INVALID, "null argument");
  if (n < 0) return fail(DDLO_E_INVALID, "bad n");
  *out = nullptr;
  DDLO_TRY(use_device(rt));
  ddlo_covs* v = nullptr;
  DDLO_TRY(covs_new(rt, n, &v));
  if (n > 0) {
    double* d_m = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&d_m), (size_t)n * 16 * sizeof(double), rt->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_m, mat4x4, (size_t)n * 16 * sizeof(double), cudaMemcpyHostToDevice, rt->stream);
    if (e == cudaSuccess) {
      k_cov_pack<<<(n + 255) / 256, 256, 0, rt->stream>>>(d_m, n, v->c);
      rt->launches += 1;
      e = cudaGetLastError();
    }
    if (d_m) cudaFreeAsync(d_m, rt->stream);
    if (e != cudaSuccess) {
      covs_free(v);
      return fail(DDLO_E_CUDA, std::string("covariance upload: ") + cudaGetErrorString(e));
    }
  }
  *out = v;
  return DDLO_OK;
}

int ddlo_covs_to_host(ddlo_covs* v, double* mat4x4_out) {
  if (!v || (v->n > 0 && !mat4x4_out)) return fail(DDLO_E_INVALID, "null argument");
  if (v->n == 0) return DDLO_OK;
  ddlo_runtime* rt = v->rt;
  DDLO_TRY(use_device(rt));
  double* d_m = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_m), (size_t)v->n * 16 * sizeof(double), rt->stream));
  k_cov_unpack<<<(v->n + 255) / 256, 256, 0, rt->stream>>>(v->c, v->n, d_m);
  rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  DDLO_CUDA(cudaMemcpyAsync(mat4x4_out, d_m, (size_t)v->n * 16 * sizeof(double), cudaMemcpyDeviceToHost, rt->stream));
  DDLO_CUDA(cudaFreeAsync(d_m, rt->stream));
  DDLO_CUDA(cudaStreamSynchronize(rt->stream));
  return DDLO_OK;
}

int ddlo_covs_size(const ddlo_covs* v, int* n) {
  if (!v || !n) return fail(DDLO_E_INVALID, "null argument");
  *n = v->n;
  return DDLO_OK;
}
int ddlo_covs_retain(ddlo_covs* v) {
  if (!v) return fail(DDLO_E_INVALID, "covs is null");
  v->refs.fetch_add(1);
  return DDLO_OK;
}
int ddlo_covs_release(ddlo_covs* v) {
  if (!v) return DDLO_OK;
  if (v->refs.fetch_sub(1) == 1) covs_free(v);
  return DDLO_OK;
}

// The same for a covariance vector; `target` (optional) is the cloud it will be the target covariances of: the
// Morton-ordered copy the align kernel reads is prepared here, once, for all engines that share the pair.
int ddlo_covs_share(ddlo_covs* v, ddlo_cloud* target) {
  if (!v) return fail(DDLO_E_INVALID, "covs is null");
  if (target && (target->rt->device != v->rt->device || target->n != v->n)) return fail(DDLO_E_SIZE, "covs_share: cloud of another device or size");
  DDLO_TRY(use_device(v->rt));
  if (target) {
    if (!target->has_index) return fail(DDLO_E_NOT_READY, "covs_share: the target cloud has no index yet");
    if (target->rt != v->rt) DDLO_CUDA(cudaStreamSynchronize(target->rt->stream));
    DDLO_TRY(ensure_sorted_covs(v, target, v->rt->stream));
  }
  DDLO_CUDA(cudaStreamSynchronize(v->rt->stream));
  v->shared = true;
  return DDLO_OK;
}

int ddlo_covs_concat(ddlo_runtime* rt, ddlo_covs* const* parts, int m, ddlo_covs** out) {
  if (!rt || !out || (m > 0 && !parts)) return fail(DDLO_E_INVALID, "null argument");
  *out = nullptr;
  DDLO_TRY(use_device(rt));
  long long total = 0;
  for (int i = 0; i < m; ++i) {
    if (!parts[i] || parts[i]->rt != rt) return fail(DDLO_E_INVALID, "concat: null part or part of another runtime");
    total += parts[i]->n;
  }
  if (total > std::numeric_limits<int>::max()) return fail(DDLO_E_UNSUPPORTED, "concat: too many points");
  ddlo_covs* r = nullptr;
  DDLO_TRY(covs_new(rt, (int)total, &r));
  std::vector<const void*> src(m);
  std::vector<size_t> words(m);
  for (int i = 0; i < m; ++i) src[i] = parts[i]->c, words[i] = (size_t)parts[i]->n * 3;  // a covariance is 48 bytes = three words
  const int rc = concat_words(rt, src.data(), words.data(), m, r->c);
  if (rc != DDLO_OK) {
    covs_free(r);
    return rc;
  }
  *out = r;
  return DDLO_OK;
}

// ---- engine ---------------------------------------------------------------------------------------------
int ddlo_params_default(ddlo_params* p) {
  if (!p) return fail(DDLO_E_INVALID, "params is null");
  std::memset(p, 0, sizeof(*p));
  p->k_correspondences = 20;
  p->regularization_method = DDLO_REG_PLANE;
  p->max_iterations = 64;
  p->optimizer = DDLO_OPT_LEVENBERG_MARQUARDT;
  p->lm_max_iterations = 10;
  p->max_correspondence_distance = (double)std::numeric_limits<float>::max();
  p->transformation_epsilon = 5e-4;
  p->rotation_epsilon = 2e-3;
  p->lm_init_lambda_factor = 1e-9;
  return DDLO_OK;
}

int ddlo_gicp_create(ddlo_runtime* rt, ddlo_gicp** out) {
  if (!rt || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = nullptr;
  DDLO_TRY(use_device(rt));
  ddlo_gicp* g = new (std::nothrow) ddlo_gicp();
  if (!g) return fail(DDLO_E_INVALID, "out of host memory");
  g->rt = rt;
  ddlo_params_default(&g->p);
  g->partial_stride = std::max(rt->max_coop_blocks_align, 64);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&g->partials), (size_t)2 * kNumSums * g->partial_stride * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&g->d_out), sizeof(AlignOut));
  if (e == cudaSuccess) runtime_retain(rt);
  if (e != cudaSuccess) {
    if (g->partials) cudaFree(g->partials);
    if (g->d_out) cudaFree(g->d_out);
    delete g;
    return fail(DDLO_E_CUDA, std::string("cudaMalloc(engine): ") + cudaGetErrorString(e));
  }
  *out = g;
  return DDLO_OK;
}

int ddlo_gicp_destroy(ddlo_gicp* g) {
  if (!g) return DDLO_OK;
  cudaSetDevice(g->rt->device);
  cudaStreamSynchronize(g->rt->stream);
  set_cloud(g->src, nullptr);
  set_cloud(g->tgt, nullptr);
  set_covs(g->src_cov, nullptr);
  set_covs(g->tgt_cov, nullptr);
  if (g->corr) cudaFreeAsync(g->corr, g->rt->stream);
  if (g->nn_seed) cudaFreeAsync(g->nn_seed, g->rt->stream);
  if (g->sqd) cudaFreeAsync(g->sqd, g->rt->stream);
  if (g->mahal) cudaFreeAsync(g->mahal, g->rt->stream);
  if (g->partials) cudaFree(g->partials);
  if (g->d_out) cudaFree(g->d_out);
  if (g->d_prof) cudaFree(g->d_prof);
  runtime_release(g->rt);
  delete g;
  return DDLO_OK;
}

int ddlo_gicp_set_params(ddlo_gicp* g, const ddlo_params* p) {
  if (!g || !p) return fail(DDLO_E_INVALID, "null argument");
  if (p->k_correspondences < 1) return fail(DDLO_E_INVALID, "k_correspondences must be >= 1");
  if (p->regularization_method < DDLO_REG_NONE || p->regularization_method > DDLO_REG_FROBENIUS)
    return fail(DDLO_E_INVALID, "unknown regularization method");
  if (p->optimizer != DDLO_OPT_GAUSS_NEWTON && p->optimizer != DDLO_OPT_LEVENBERG_MARQUARDT)
    return fail(DDLO_E_INVALID, "unknown optimizer");
  if (p->max_iterations < 0 || p->lm_max_iterations < 0) return fail(DDLO_E_INVALID, "negative iteration limit");
  g->p = *p;
  return DDLO_OK;
}
int ddlo_gicp_get_params(const ddlo_gicp* g, ddlo_params* p) {
  if (!g || !p) return fail(DDLO_E_INVALID, "null argument");
  *p = g->p;
  return DDLO_OK;
}

// a handle of another runtime is accepted only once it has been shared (complete, synchronised, read-only)
static int check_foreign(const ddlo_gicp* g, const ddlo_runtime* owner, bool shared, const char* what) {
  if (owner == g->rt) return DDLO_OK;
  if (owner->device != g->rt->device) return fail(DDLO_E_INVALID, std::string(what) + " lives on another device");
  if (!shared) return fail(DDLO_E_INVALID, std::string(what) + " belongs to another runtime (ddlo_cloud_share / ddlo_covs_share make it usable from other streams)");
  return DDLO_OK;
}

int ddlo_gicp_set_input_source(ddlo_gicp* g, ddlo_cloud* c, int build) {
  if (!g || !c) return fail(DDLO_E_INVALID, "null argument");
  DDLO_TRY(check_foreign(g, c->rt, c->shared, "cloud"));
  if (g->src == c) return DDLO_OK;  // if (input_ == cloud) return;
  DDLO_TRY(use_device(g->rt));
  set_cloud(g->src, c);
  g->corr_n = 0;  // the stored correspondences index the previous source
  if (build) {
    if (c->n > 0) DDLO_TRY(build_index(c));
    set_covs(g->src_cov, nullptr);
  }
  return DDLO_OK;
}

int ddlo_gicp_set_input_target(ddlo_gicp* g, ddlo_cloud* c) {
  if (!g || !c) return fail(DDLO_E_INVALID, "null argument");
  DDLO_TRY(check_foreign(g, c->rt, c->shared, "cloud"));
  if (g->tgt == c) return DDLO_OK;
  DDLO_TRY(use_device(g->rt));
  set_cloud(g->tgt, c);
  g->corr_n = 0;  // the stored correspondences index points of the previous target
  if (c->n > 0) DDLO_TRY(build_index(c));
  set_covs(g->tgt_cov, nullptr);
  return DDLO_OK;
}

int ddlo_gicp_clear_source(ddlo_gicp* g) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  set_cloud(g->src, nullptr);
  set_covs(g->src_cov, nullptr);
  g->corr_n = 0;
  return DDLO_OK;
}
int ddlo_gicp_clear_target(ddlo_gicp* g) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  set_cloud(g->tgt, nullptr);
  set_covs(g->tgt_cov, nullptr);
  g->corr_n = 0;
  return DDLO_OK;
}

int ddlo_gicp_set_source_covariances(ddlo_gicp* g, ddlo_covs* v) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  if (v) DDLO_TRY(check_foreign(g, v->rt, v->shared, "covariance vector"));
  set_covs(g->src_cov, v);
  return DDLO_OK;
}
int ddlo_gicp_set_target_covariances(ddlo_gicp* g, ddlo_covs* v) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  if (v) DDLO_TRY(check_foreign(g, v->rt, v->shared, "covariance vector"));
  set_covs(g->tgt_cov, v);
  return DDLO_OK;
}
int ddlo_gicp_get_source_covariances(ddlo_gicp* g, ddlo_covs** out) {
  if (!g || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = g->src_cov;
  if (*out) (*out)->refs.fetch_add(1);
  return DDLO_OK;
}
int ddlo_gicp_get_target_covariances(ddlo_gicp* g, ddlo_covs** out) {
  if (!g || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = g->tgt_cov;
  if (*out) (*out)->refs.fetch_add(1);
  return DDLO_OK;
}
int ddlo_gicp_get_input_source(ddlo_gicp* g, ddlo_cloud** out) {
  if (!g || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = g->src;
  if (*out) (*out)->refs.fetch_add(1);
  return DDLO_OK;
}
int ddlo_gicp_get_input_target(ddlo_gicp* g, ddlo_cloud** out) {
  if (!g || !out) return fail(DDLO_E_INVALID, "null argument");
  *out = g->tgt;
  if (*out) (*out)->refs.fetch_add(1);
  return DDLO_OK;
}

static int calc_covs(ddlo_gicp* g, ddlo_cloud* c, ddlo_covs*& slot) {
  if (!c) return fail(DDLO_E_NOT_READY, "calculate covariances: no cloud set");
  ddlo_covs* v = nullptr;
  DDLO_TRY(ddlo_covs_compute(c, g->p.k_correspondences, g->p.regularization_method, &v));
  // a shared cloud of another runtime: its covariances are computed on the owner's stream; this engine's stream must
  // not run ahead of them (rare path: callers that share a target normally share its covariances too)
  if (c->rt != g->rt) DDLO_CUDA(cudaStreamSynchronize(c->rt->stream));
  set_covs(slot, v);
  ddlo_covs_release(v);
  return DDLO_OK;
}
int ddlo_gicp_calculate_source_covariances(ddlo_gicp* g) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  return calc_covs(g, g->src, g->src_cov);
}
int ddlo_gicp_calculate_target_covariances(ddlo_gicp* g) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  return calc_covs(g, g->tgt, g->tgt_cov);
}

int ddlo_gicp_swap_source_and_target(ddlo_gicp* g) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  std::swap(g->src, g->tgt);          // input_.swap(target_) and the kd-trees with them
  std::swap(g->src_cov, g->tgt_cov);  // source_covs_.swap(target_covs_)
  g->corr_n = 0;                      // correspondences_.clear(); sq_distances_.clear();
  return DDLO_OK;
}

}  // extern "C"

// covs_sorted[p] = covs[original index of the p-th point in Morton order]
__global__ void __launch_bounds__(256) k_permute_covs(const float4* __restrict__ spts, int n, const double* __restrict__ covs,
                                                      double* __restrict__ covs_sorted) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int o = __float_as_int(spts[p].w);
  const double2* in = reinterpret_cast<const double2*>(covs + (size_t)o * kCovStride);
  double2* out = reinterpret_cast<double2*>(covs_sorted + (size_t)p * kCovStride);
  out[0] = in[0];
  out[1] = in[1];
  out[2] = in[2];
}

// The kernels gather the matched target covariance by the match's position in the Morton order, so
// that neighbouring source points (whose matches are neighbours too) touch the same DRAM pages.
// The permuted copy lives on the covariance handle and is rebuilt only when it is paired with another
// cloud; engines that share a target (S2S -> S2M hand-over, the lanes of a batch) share the copy.
int ddlo::ensure_sorted_covs(ddlo_covs* v, ddlo_cloud* cloud, cudaStream_t st) {
  if (v->sorted && v->sorted_for == cloud) return DDLO_OK;
  if (v->shared) return fail(DDLO_E_NOT_READY, "shared covariances were prepared for another target cloud (ddlo_covs_share)");
  const size_t n = (size_t)cloud->n;
  if (v->sorted) {
    DDLO_CUDA(cudaFreeAsync(v->sorted, st));
    v->sorted = nullptr;
  }
  cloud_set(v->sorted_for, nullptr);
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&v->sorted), n * kCovStride * sizeof(double), st));
  k_permute_covs<<<(int)((n + 255) / 256), 256, 0, st>>>(cloud->spts, (int)n, v->c, v->sorted);
  v->rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  cloud_set(v->sorted_for, cloud);  // retained: the key can not be recycled while cached
  return DDLO_OK;
}

static int ensure_workspace(ddlo_gicp* g, int ns) {
  if (g->ws_n >= ns) return DDLO_OK;
  // stream-ordered: the frees run behind the kernels that still use the old buffers, the allocations come
  // from the pool warmed at runtime creation (no cudaMalloc, no synchronisation, in the middle of a frame)
  cudaStream_t st = g->rt->stream;
  if (g->corr) cudaFreeAsync(g->corr, st);
  if (g->nn_seed) cudaFreeAsync(g->nn_seed, st);
  if (g->sqd) cudaFreeAsync(g->sqd, st);
  if (g->mahal) cudaFreeAsync(g->mahal, st);
  g->corr = nullptr, g->nn_seed = nullptr, g->sqd = nullptr, g->mahal = nullptr, g->ws_n = 0;
  const size_t cap = (size_t)ns + ns / 4 + 256;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&g->corr), cap * sizeof(int), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&g->nn_seed), cap * sizeof(int2), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&g->sqd), cap * sizeof(float), st));
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&g->mahal), cap * kCovStride * sizeof(double), st));
  g->ws_n = (int)cap;
  return DDLO_OK;
}

// everything align/linearize need, or the reason they can not run
static int align_blocks(const ddlo_gicp* g);

// guess16_pass0 != nullptr (align() proper, not the stepwise hooks or the batched waves): when covariances have to be
// computed first, the correspondence search of the first linearize - it needs the guess, the source points and the
// target index, not the covariances - runs as a kernel of its own on a side stream beside the covariance kernels.
static int prepare(ddlo_gicp* g, bool compute_missing_covs, int* covs_computed, GicpArgs* a, const float* guess16_pass0 = nullptr) {
  if (!g->src || !g->tgt) return fail(DDLO_E_NOT_READY, "input source and target must both be set");
  if (g->src->n <= 0 || g->tgt->n <= 0) return fail(DDLO_E_EMPTY, "source or target cloud is empty");
  DDLO_TRY(use_device(g->rt));
  DDLO_TRY(build_index(g->tgt));
  DDLO_TRY(ensure_workspace(g, g->src->n));
  const bool need_src = !g->src_cov || g->src_cov->n != g->src->n, need_tgt = !g->tgt_cov || g->tgt_cov->n != g->tgt->n;
  bool pass0 = false;
  // opt-in (DDLO_PASS0_OVERLAP=1): measured on B200 at -2 us of 480 for the C2 step - the covariance search is issue
  // bound, so the early search only takes its issue slots (profiles/README.md)
  static const bool pass0_overlap = std::getenv("DDLO_PASS0_OVERLAP") != nullptr;
  if (guess16_pass0 && compute_missing_covs && (need_src || need_tgt) && pass0_overlap && g->rt->side2 && g->p.max_iterations > 0 &&
      g->src->rt == g->rt && g->tgt->rt == g->rt) {
    ddlo_runtime* rt = g->rt;
    GicpArgs sa;
    std::memset(&sa, 0, sizeof(sa));
    sa.tgt = g->tgt->view;
    sa.src_pts = g->src->pts;
    sa.ns = g->src->n;
    sa.corr = g->corr;
    sa.nn_seed = g->nn_seed;
    sa.sqd = g->sqd;
    sa.thr2 = g->p.max_correspondence_distance * g->p.max_correspondence_distance;
    std::memcpy(sa.guess, guess16_pass0, sizeof(sa.guess));
    DDLO_CUDA(cudaEventRecord(rt->ev_fork2, rt->stream));  // behind the target index, the source upload and the workspace
    DDLO_CUDA(cudaStreamWaitEvent(rt->side2, rt->ev_fork2, 0));
    DDLO_TRY(launch_search_pass0(rt, rt->side2, sa, align_blocks(g)));
    DDLO_CUDA(cudaEventRecord(rt->ev_join2, rt->side2));
    pass0 = true;
  }
  if (!g->src_cov || g->src_cov->n != g->src->n) {
    if (!compute_missing_covs) return fail(g->src_cov ? DDLO_E_SIZE : DDLO_E_NOT_READY, "source covariances missing or of the wrong size");
    DDLO_TRY(calc_covs(g, g->src, g->src_cov));
    if (covs_computed) *covs_computed = 1;
  }
  if (!g->tgt_cov || g->tgt_cov->n != g->tgt->n) {
    if (!compute_missing_covs) return fail(g->tgt_cov ? DDLO_E_SIZE : DDLO_E_NOT_READY, "target covariances missing or of the wrong size");
    DDLO_TRY(calc_covs(g, g->tgt, g->tgt_cov));
    if (covs_computed) *covs_computed = 1;
  }
  DDLO_TRY(ensure_sorted_covs(g->tgt_cov, g->tgt, g->rt->stream));
  if (pass0) DDLO_CUDA(cudaStreamWaitEvent(g->rt->stream, g->rt->ev_join2, 0));
  a->pass0_done = pass0 ? 1 : 0;
  a->tgt = g->tgt->view;
  a->src_pts = g->src->pts;
  a->src_lattice = g->src->has_index ? g->src->lattice : nullptr;
  a->src_cov = g->src_cov->c;
  a->tgt_pts = g->tgt->pts;
  a->tgt_cov = g->tgt_cov->sorted;
  a->ns = g->src->n;
  a->corr = g->corr;
  a->nn_seed = g->nn_seed;
  a->sqd = g->sqd;
  a->mahal = g->mahal;
  a->partials = g->partials;
  a->partial_stride = g->partial_stride;
  a->max_iterations = g->p.max_iterations;
  a->optimizer = g->p.optimizer;
  a->lm_max_iterations = g->p.lm_max_iterations;
  a->thr2 = g->p.max_correspondence_distance * g->p.max_correspondence_distance;
  a->trans_eps = g->p.transformation_epsilon;
  a->rot_eps = g->p.rotation_epsilon;
  a->lm_init_lambda_factor = g->p.lm_init_lambda_factor;
  a->out = g->d_out;
  a->blk_times = nullptr;
  a->stamps = nullptr;
  if (g->profile) {
    const size_t words = (size_t)8 * g->partial_stride * 8 + kStampCap + 1;
    if (!g->d_prof) DDLO_CUDA(cudaMalloc(reinterpret_cast<void**>(&g->d_prof), words * sizeof(unsigned long long)));
    DDLO_CUDA(cudaMemsetAsync(g->d_prof, 0, words * sizeof(unsigned long long), g->rt->stream));
    a->blk_times = g->d_prof;
    a->stamps = g->d_prof + (size_t)8 * g->partial_stride * 8;
  }
  a->dbg_visits = nullptr;
#ifdef DDLO_VISIT_STATS
  if (g->d_dbg_n < g->src->n) {
    if (g->d_dbg) cudaFree(g->d_dbg);
    DDLO_CUDA(cudaMalloc(reinterpret_cast<void**>(&g->d_dbg), (size_t)4 * g->src->n * sizeof(int4)));
    g->d_dbg_n = g->src->n;
  }
  a->dbg_visits = g->d_dbg;
#endif
  return DDLO_OK;
}

static int align_blocks(const ddlo_gicp* g) { return gicp_blocks_for(g->src->n, g->rt->align_blocks_limit); }

// the argument record of one align (missing covariances are computed first, on the engine's stream) and the number
// of blocks / chunks it runs with
int ddlo::prepare_align(ddlo_gicp* g, const float* guess16, int* covs_computed, GicpArgs* a, int* nblocks, bool single_launch) {
  static const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  DDLO_TRY(prepare(g, true, covs_computed, a, single_launch ? (guess16 ? guess16 : I16) : nullptr));
  std::memcpy(a->guess, guess16 ? guess16 : I16, sizeof(a->guess));
  std::memset(a->T_step, 0, sizeof(a->T_step));
  *nblocks = align_blocks(g);
  g->corr_n = g->p.max_iterations > 0 ? g->src->n : g->corr_n;
  return DDLO_OK;
}

int ddlo::enqueue_align(ddlo_gicp* g, const float* guess16, int* covs_computed) {
  GicpArgs a;
  int blocks = 0;
  DDLO_TRY(prepare_align(g, guess16, covs_computed, &a, &blocks, true));
  DDLO_TRY(launch_align(g->rt, a, blocks));
  return DDLO_OK;
}

void ddlo::fill_result(const AlignOut* o, int covs_computed, ddlo_align_result* r) {
  std::memcpy(r->final_transformation, o->final_transformation, sizeof(r->final_transformation));
  std::memcpy(r->final_hessian, o->final_hessian, sizeof(r->final_hessian));
  r->flags = o->flags | (covs_computed ? DDLO_FLAG_COVS_COMPUTED : 0);
  r->nr_iterations = o->nr_iterations;
  r->n_linearize = o->n_linearize;
  r->n_compute_error = o->n_compute_error;
  r->final_error = o->final_error;
  r->lm_lambda = o->lm_lambda;
}

extern "C" {

int ddlo_gicp_align_async(ddlo_gicp* g, const float* guess16) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  g->pending_covs_computed = 0;
  DDLO_TRY(enqueue_align(g, guess16, &g->pending_covs_computed));
  g->align_pending = true;
  return DDLO_OK;
}

int ddlo_gicp_align_finish(ddlo_gicp* g, ddlo_align_result* result) {
  if (!g || !result) return fail(DDLO_E_INVALID, "null argument");
  if (!g->align_pending) return fail(DDLO_E_NOT_READY, "align_finish without align_async");
  g->align_pending = false;
  ddlo_runtime* rt = g->rt;
  DDLO_TRY(use_device(rt));
  DDLO_TRY(ensure_pinned(rt, sizeof(AlignOut)));
  DDLO_CUDA(cudaMemcpyAsync(rt->h_pinned, g->d_out, sizeof(AlignOut), cudaMemcpyDeviceToHost, rt->stream));
  DDLO_CUDA(cudaStreamSynchronize(rt->stream));
  fill_result(static_cast<const AlignOut*>(rt->h_pinned), g->pending_covs_computed, result);
  std::memcpy(g->last_T, result->final_transformation, sizeof(g->last_T));
  g->has_last_T = true;
  if (result->flags & DDLO_FLAG_NONFINITE) return fail(DDLO_E_NONFINITE, "align: an input cloud holds NaN or Inf coordinates");
  return DDLO_OK;
}

int ddlo_gicp_align(ddlo_gicp* g, const float* guess16, ddlo_align_result* result) {
  if (!g || !result) return fail(DDLO_E_INVALID, "null argument");
  DDLO_TRY(ddlo_gicp_align_async(g, guess16));
  return ddlo_gicp_align_finish(g, result);
}

int ddlo_gicp_align_batch(ddlo_gicp* const* engines, int m, const float* guesses16, ddlo_align_result* results) {
  if (m < 0 || (m > 0 && (!engines || !results))) return fail(DDLO_E_INVALID, "null argument");
  if (m == 0) return DDLO_OK;
  ddlo_runtime* rt = engines[0] ? engines[0]->rt : nullptr;
  if (!rt) return fail(DDLO_E_INVALID, "null engine");
  for (int i = 0; i < m; ++i)
    if (!engines[i] || engines[i]->rt != rt) return fail(DDLO_E_INVALID, "batch engines must share one runtime");
  DDLO_TRY(use_device(rt));
  DDLO_TRY(ensure_pinned(rt, sizeof(AlignOut) * (size_t)m));
  std::vector<int> covs_computed(m, 0);
  AlignOut* h = static_cast<AlignOut*>(rt->h_pinned);
  int rc = DDLO_OK, done = 0;
  for (; done < m && rc == DDLO_OK; ++done) {
    rc = enqueue_align(engines[done], guesses16 ? guesses16 + 16 * (size_t)done : nullptr, &covs_computed[done]);
    if (rc == DDLO_OK && cudaMemcpyAsync(h + done, engines[done]->d_out, sizeof(AlignOut), cudaMemcpyDeviceToHost, rt->stream) != cudaSuccess)
      rc = fail(DDLO_E_CUDA, "align_batch: result copy failed");
    if (rc != DDLO_OK) break;
  }
  // whatever was enqueued completes before this call returns, also on failure
  if (cudaStreamSynchronize(rt->stream) != cudaSuccess && rc == DDLO_OK) rc = fail(DDLO_E_CUDA, "align_batch: synchronisation failed");
  for (int i = 0; i < done; ++i) {
    fill_result(h + i, covs_computed[i], results + i);
    std::memcpy(engines[i]->last_T, results[i].final_transformation, sizeof(engines[i]->last_T));
    engines[i]->has_last_T = true;
    if (rc == DDLO_OK && (results[i].flags & DDLO_FLAG_NONFINITE)) rc = fail(DDLO_E_NONFINITE, "align_batch: an input cloud holds NaN or Inf coordinates");
  }
  return rc;
}

int ddlo_gicp_aligned_cloud(ddlo_gicp* g, ddlo_cloud** out) {
  if (!g || !out) return fail(DDLO_E_INVALID, "null argument");
  if (!g->src || !g->has_last_T) return fail(DDLO_E_NOT_READY, "aligned cloud requested before align");
  return ddlo_cloud_transform(g->src, g->last_T, out);
}

static int read_out(ddlo_gicp* g, AlignOut* host) {
  ddlo_runtime* rt = g->rt;
  DDLO_TRY(ensure_pinned(rt, sizeof(AlignOut)));
  DDLO_CUDA(cudaMemcpyAsync(rt->h_pinned, g->d_out, sizeof(AlignOut), cudaMemcpyDeviceToHost, rt->stream));
  DDLO_CUDA(cudaStreamSynchronize(rt->stream));
  std::memcpy(host, rt->h_pinned, sizeof(AlignOut));
  return DDLO_OK;
}

int ddlo_gicp_linearize(ddlo_gicp* g, const double* T16, double* H36, double* b6, double* error) {
  if (!g || !T16) return fail(DDLO_E_INVALID, "null argument");
  GicpArgs a;
  DDLO_TRY(prepare(g, false, nullptr, &a));
  std::memset(a.guess, 0, sizeof(a.guess));
  std::memcpy(a.T_step, T16, sizeof(a.T_step));
  DDLO_TRY(launch_linearize_step(g->rt, a, align_blocks(g)));
  g->corr_n = a.ns;
  AlignOut o;
  DDLO_TRY(read_out(g, &o));
  const double* t = o.sums;
  if (H36) {
    double H[36];
    // same unpacking as the device (row-major == column-major for a symmetric matrix)
    H[0] = t[0], H[1] = t[1], H[2] = t[2], H[7] = t[3], H[8] = t[4], H[14] = t[5];
    H[6] = t[1], H[12] = t[2], H[13] = t[4];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        H[6 * r + 3 + c] = t[6 + 3 * r + c];
        H[6 * (3 + c) + r] = t[6 + 3 * r + c];
      }
    H[21] = t[15], H[22] = t[16], H[23] = t[17], H[28] = t[18], H[29] = t[19], H[35] = t[20];
    H[27] = t[16], H[33] = t[17], H[34] = t[19];
    std::memcpy(H36, H, sizeof(H));
  }
  if (b6)
    for (int r = 0; r < 6; ++r) b6[r] = t[21 + r];
  if (error) *error = t[27];
  return DDLO_OK;
}

int ddlo_gicp_compute_error(ddlo_gicp* g, const double* T16, double* error) {
  if (!g || !T16 || !error) return fail(DDLO_E_INVALID, "null argument");
  if (!g->src || g->corr_n != g->src->n) return fail(DDLO_E_NOT_READY, "compute_error needs the correspondences of a previous linearize");
  GicpArgs a;
  DDLO_TRY(prepare(g, false, nullptr, &a));
  std::memset(a.guess, 0, sizeof(a.guess));
  std::memcpy(a.T_step, T16, sizeof(a.T_step));
  DDLO_TRY(launch_error_step(g->rt, a, align_blocks(g)));
  AlignOut o;
  DDLO_TRY(read_out(g, &o));
  *error = o.sums[0];
  return DDLO_OK;
}

int ddlo_gicp_get_correspondences(ddlo_gicp* g, int* correspondences, float* sq_distances, int capacity) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  if (!g->src || g->corr_n != g->src->n) return fail(DDLO_E_NOT_READY, "no correspondences: run align or linearize first");
  if (capacity < g->corr_n) return fail(DDLO_E_SIZE, "output capacity is smaller than the source cloud");
  ddlo_runtime* rt = g->rt;
  DDLO_TRY(use_device(rt));
  if (correspondences) DDLO_CUDA(cudaMemcpyAsync(correspondences, g->corr, (size_t)g->corr_n * sizeof(int), cudaMemcpyDeviceToHost, rt->stream));
  if (sq_distances) DDLO_CUDA(cudaMemcpyAsync(sq_distances, g->sqd, (size_t)g->corr_n * sizeof(float), cudaMemcpyDeviceToHost, rt->stream));
  DDLO_CUDA(cudaStreamSynchronize(rt->stream));
  return DDLO_OK;
}

int ddlo_gicp_get_mahalanobis(ddlo_gicp* g, double* mat4x4_out, int capacity) {
  if (!g || !mat4x4_out) return fail(DDLO_E_INVALID, "null argument");
  if (!g->src || g->corr_n != g->src->n) return fail(DDLO_E_NOT_READY, "no Mahalanobis matrices: run align or linearize first");
  if (capacity < g->corr_n) return fail(DDLO_E_SIZE, "output capacity is smaller than the source cloud");
  ddlo_runtime* rt = g->rt;
  DDLO_TRY(use_device(rt));
  const int n = g->corr_n;
  double* d_m = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_m), (size_t)n * 16 * sizeof(double), rt->stream));
  k_cov_unpack<<<(n + 255) / 256, 256, 0, rt->stream>>>(g->mahal, n, d_m);
  rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  DDLO_CUDA(cudaMemcpyAsync(mat4x4_out, d_m, (size_t)n * 16 * sizeof(double), cudaMemcpyDeviceToHost, rt->stream));
  DDLO_CUDA(cudaFreeAsync(d_m, rt->stream));
  DDLO_CUDA(cudaStreamSynchronize(rt->stream));
  return DDLO_OK;
}

int ddlo_gicp_get_residuals_async(ddlo_gicp* g, double* out, int capacity) {
  if (!g || !out) return fail(DDLO_E_INVALID, "null argument");
  if (!g->src || g->corr_n != g->src->n) return fail(DDLO_E_NOT_READY, "no residuals: run align first");
  if (capacity < g->corr_n) return fail(DDLO_E_SIZE, "output capacity is smaller than the source cloud");
  ddlo_runtime* rt = g->rt;
  DDLO_TRY(use_device(rt));
  const int n = g->corr_n;
  double* d_r = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_r), (size_t)n * sizeof(double), rt->stream));
  k_sqrt_f2d<<<(n + 255) / 256, 256, 0, rt->stream>>>(g->sqd, n, d_r);
  rt->launches += 1;
  DDLO_CUDA(cudaGetLastError());
  DDLO_CUDA(cudaMemcpyAsync(out, d_r, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, rt->stream));
  DDLO_CUDA(cudaFreeAsync(d_r, rt->stream));
  return DDLO_OK;
}

int ddlo_gicp_get_residuals(ddlo_gicp* g, double* out, int capacity) {
  DDLO_TRY(ddlo_gicp_get_residuals_async(g, out, capacity));
  DDLO_CUDA(cudaStreamSynchronize(g->rt->stream));
  return DDLO_OK;
}

int ddlo_gicp_residual_image(ddlo_gicp* g, int width, int height, double angle_min, double angle_max, float* out_xyzi) {
  if (!g || !out_xyzi) return fail(DDLO_E_INVALID, "null argument");
  if (width <= 0 || height <= 0 || (long long)width * height > (1 << 26) || !(angle_max > angle_min)) return fail(DDLO_E_INVALID, "bad image geometry");
  if (!g->src || g->corr_n != g->src->n) return fail(DDLO_E_NOT_READY, "no residuals: run align first");
  ddlo_runtime* rt = g->rt;
  DDLO_TRY(use_device(rt));
  const size_t cells = (size_t)width * height;
  float4* d_img = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_img), cells * sizeof(float4), rt->stream));
  int rc = residual_image_device(rt, g->src->pts, g->sqd, g->corr_n, width, height, angle_min, angle_max, d_img);
  if (rc == DDLO_OK) {
    cudaError_t e = cudaMemcpyAsync(out_xyzi, d_img, cells * sizeof(float4), cudaMemcpyDeviceToHost, rt->stream);
    if (e != cudaSuccess) rc = fail(DDLO_E_CUDA, cudaGetErrorString(e));
  }
  cudaFreeAsync(d_img, rt->stream);
  if (rc != DDLO_OK) return rc;
  DDLO_CUDA(cudaStreamSynchronize(rt->stream));
  return DDLO_OK;
}

// residuals: host plane, or (engine != null) the residual image of the engine's last align built on the device
static int segment_scan_impl(ddlo_runtime* rt, const ddlo_segmentation_params* params, const float* scan_t, int stride_bytes, const float* T16,
                             const float* residuals, ddlo_gicp* engine, double angle_min, double angle_max, int* label_mat, float* range_mat,
                             signed char* ground_mat, double* avg_residuals, int avg_capacity, int* label_count, float* device_ms) {
  if (!rt || !params || !scan_t || !T16 || !label_count) return fail(DDLO_E_INVALID, "null argument");
  const ddlo_segmentation_params& p = *params;
  if (p.rows < 2 || p.cols < 1 || (long long)p.rows * p.cols > (1 << 24)) return fail(DDLO_E_INVALID, "bad range image geometry");
  if (stride_bytes < 12 || stride_bytes % 4 != 0) return fail(DDLO_E_INVALID, "stride must be a multiple of 4, at least 12 bytes");
  if (p.ground_rows < 0 || p.ground_rows >= p.rows) return fail(DDLO_E_INVALID, "ground_rows must be below rows (the reference indexes row -1 otherwise)");
  if (p.window_col_max >= p.cols)
    return fail(DDLO_E_UNSUPPORTED, "window_col_max >= cols: the reference's one-directional column wrap (detection.cpp:590-599) is not supported");
  if (avg_residuals && avg_capacity < 0) return fail(DDLO_E_INVALID, "negative capacity");
  DDLO_TRY(use_device(rt));
  cudaStream_t st = rt->stream;
  const size_t HW = (size_t)p.rows * p.cols;
  const int stride = stride_bytes / 4;
  // stream-ordered pool buffers, returned to the pool on every exit path
  struct PoolBuf {
    void* p = nullptr;
    cudaStream_t st;
    explicit PoolBuf(cudaStream_t s) : st(s) {}
    ~PoolBuf() {
      if (p) cudaFreeAsync(p, st);
    }
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, st); }
  };
  PoolBuf b_scan(st), b_res(st), b_range(st), b_label(st), b_ground(st), b_avg(st);
  DDLO_CUDA(b_scan.alloc(HW * stride_bytes));
  DDLO_CUDA(b_range.alloc(HW * 4));
  DDLO_CUDA(b_label.alloc(HW * 4 + 16));
  DDLO_CUDA(b_ground.alloc(HW));
  DDLO_CUDA(b_avg.alloc((HW + 1) * 8));
  if (residuals) DDLO_CUDA(b_res.alloc(HW * 4));
  if (engine) DDLO_CUDA(b_res.alloc(HW * sizeof(float4)));
  float *d_scan = static_cast<float*>(b_scan.p), *d_res = static_cast<float*>(b_res.p), *d_range = static_cast<float*>(b_range.p);
  int* d_label = static_cast<int*>(b_label.p);
  int* d_count = d_label + HW;
  signed char* d_ground = static_cast<signed char*>(b_ground.p);
  double* d_avg = static_cast<double*>(b_avg.p);
  DDLO_CUDA(cudaMemcpyAsync(d_scan, scan_t, HW * stride_bytes, cudaMemcpyHostToDevice, st));
  if (residuals) DDLO_CUDA(cudaMemcpyAsync(d_res, residuals, HW * 4, cudaMemcpyHostToDevice, st));
  DDLO_CUDA(cudaMemsetAsync(d_avg, 0, (HW + 1) * 8, st));
  if (device_ms) DDLO_CUDA(cudaEventRecord(rt->ev0, st));
  if (engine)  // the residual cloud of odom.cc:804-827, cols x rows cells; its intensity channel is the residual plane (:240-249)
    DDLO_TRY(residual_image_device(rt, engine->src->pts, engine->sqd, engine->corr_n, p.cols, p.rows, angle_min, angle_max,
                                   reinterpret_cast<float4*>(d_res)));
  DDLO_TRY(segment_scan_device(rt, p, T16, d_scan, stride, engine ? d_res + 3 : d_res, engine ? 4 : 1, d_label, d_range, d_ground, d_avg, d_count));
  if (device_ms) DDLO_CUDA(cudaEventRecord(rt->ev1, st));
  // One synchronisation in the common case: the segment count travels together with the first kAvgFirst averages.
  constexpr int kAvgFirst = 1024;
  const int first = avg_residuals ? std::min<long long>(std::min(avg_capacity, kAvgFirst), (long long)HW + 1) : 0;
  int count = 0;
  DDLO_CUDA(cudaMemcpyAsync(&count, d_count, 4, cudaMemcpyDeviceToHost, st));
  if (first > 0) DDLO_CUDA(cudaMemcpyAsync(avg_residuals, d_avg, (size_t)first * 8, cudaMemcpyDeviceToHost, st));
  if (label_mat) DDLO_CUDA(cudaMemcpyAsync(label_mat, d_label, HW * 4, cudaMemcpyDeviceToHost, st));
  if (range_mat) DDLO_CUDA(cudaMemcpyAsync(range_mat, d_range, HW * 4, cudaMemcpyDeviceToHost, st));
  if (ground_mat) DDLO_CUDA(cudaMemcpyAsync(ground_mat, d_ground, HW, cudaMemcpyDeviceToHost, st));
  DDLO_CUDA(cudaStreamSynchronize(st));
  const int wanted = avg_residuals ? std::min(count, avg_capacity) : 0;
  if (wanted > first) {
    DDLO_CUDA(cudaMemcpyAsync(avg_residuals + first, d_avg + first, (size_t)(wanted - first) * 8, cudaMemcpyDeviceToHost, st));
    DDLO_CUDA(cudaStreamSynchronize(st));
  }
  if (device_ms) DDLO_CUDA(cudaEventElapsedTime(device_ms, rt->ev0, rt->ev1));
  *label_count = count;
  return DDLO_OK;
}

int ddlo_segment_scan(ddlo_runtime* rt, const ddlo_segmentation_params* params, const float* scan_t, int stride_bytes, const float* T16,
                      const float* residuals, int* label_mat, float* range_mat, signed char* ground_mat, double* avg_residuals,
                      int avg_capacity, int* label_count, float* device_ms) {
  return segment_scan_impl(rt, params, scan_t, stride_bytes, T16, residuals, nullptr, 0.0, 0.0, label_mat, range_mat, ground_mat, avg_residuals,
                           avg_capacity, label_count, device_ms);
}

int ddlo_gicp_segment_scan(ddlo_gicp* g, const ddlo_segmentation_params* params, const float* scan_t, int stride_bytes, const float* T16,
                           double angle_min, double angle_max, int* label_mat, float* range_mat, signed char* ground_mat, double* avg_residuals,
                           int avg_capacity, int* label_count, float* device_ms) {
  if (!g) return fail(DDLO_E_INVALID, "null argument");
  if (!(angle_max > angle_min)) return fail(DDLO_E_INVALID, "bad image geometry");
  if (!g->src || g->corr_n != g->src->n) return fail(DDLO_E_NOT_READY, "no residuals: run align first");
  return segment_scan_impl(g->rt, params, scan_t, stride_bytes, T16, nullptr, g, angle_min, angle_max, label_mat, range_mat, ground_mat,
                           avg_residuals, avg_capacity, label_count, device_ms);
}

int ddlo_gicp_get_residual_vectors(ddlo_gicp* g, const float* T16, float* out_xyz, int capacity) {
  if (!g || !T16 || !out_xyz) return fail(DDLO_E_INVALID, "null argument");
  if (!g->src || !g->tgt || g->corr_n != g->src->n) return fail(DDLO_E_NOT_READY, "no residuals: run align first");
  if (capacity < g->corr_n) return fail(DDLO_E_SIZE, "output capacity is smaller than the source cloud");
  ddlo_runtime* rt = g->rt;
  DDLO_TRY(use_device(rt));
  const int n = g->corr_n;
  float* d_r = nullptr;
  DDLO_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_r), (size_t)n * 3 * sizeof(float), rt->stream));
  DDLO_TRY(launch_residual_vectors(rt, g->src->pts, g->tgt->pts, g->corr, n, T16, d_r));
  DDLO_CUDA(cudaMemcpyAsync(out_xyz, d_r, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, rt->stream));
  DDLO_CUDA(cudaFreeAsync(d_r, rt->stream));
  DDLO_CUDA(cudaStreamSynchronize(rt->stream));
  return DDLO_OK;
}

// debugging aids (declared in ddlo_gicp_testing.h).  Profiling is off by default: the align kernel then takes no
// timestamps and touches no profiling memory.
int ddlo_gicp_debug_enable(ddlo_gicp* g, int on) {
  if (!g) return fail(DDLO_E_INVALID, "engine is null");
  g->profile = on != 0;
  return DDLO_OK;
}

// the phase timeline block 0 recorded during the last align; entries are tag << 56 | globaltimer ns.
// Returns the number of entries written.
int ddlo_gicp_debug_timeline(ddlo_gicp* g, unsigned long long* out, int capacity) {
  if (!g || !out) return fail(DDLO_E_INVALID, "null argument");
  if (!g->profile || !g->d_prof) return 0;
  if (cudaStreamSynchronize(g->rt->stream) != cudaSuccess) return DDLO_E_CUDA;
  std::vector<unsigned long long> h(kStampCap + 1);
  if (cudaMemcpy(h.data(), g->d_prof + (size_t)8 * g->partial_stride * 8, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return DDLO_E_CUDA;
  const int n = std::min(std::min((int)h[0], kStampCap), capacity);
  for (int i = 0; i < n; ++i) out[i] = h[1 + i];
  return n;
}

// per-block phase times of the first 8 linearize passes of the last align,
// out[pass][block][8] in ns; returns the number of blocks the kernel ran with (0 when profiling is off).
int ddlo_gicp_debug_block_times(ddlo_gicp* g, unsigned long long* out, int capacity_blocks) {
  if (!g || !out || !g->src) return fail(DDLO_E_INVALID, "null argument");
  if (!g->profile || !g->d_prof) return 0;
  const int nb = align_blocks(g);
  if (capacity_blocks < nb) return fail(DDLO_E_SIZE, "capacity too small");
  DDLO_CUDA(cudaStreamSynchronize(g->rt->stream));
  for (int p = 0; p < 8; ++p)
    DDLO_CUDA(cudaMemcpy(out + (size_t)p * capacity_blocks * 8, g->d_prof + (size_t)p * nb * 8, (size_t)nb * 8 * sizeof(unsigned long long),
                         cudaMemcpyDeviceToHost));
  return nb;
}

// DDLO_VISIT_STATS builds only: per source point {node visits, leaf scans, warp steps, 0} of the first 4
// linearize passes of the last align, out[4][ns][4]; returns ns (0 in regular builds)
int ddlo_gicp_debug_visits(ddlo_gicp* g, int* out, int capacity_points) {
  if (!g || !out || !g->src) return fail(DDLO_E_INVALID, "null argument");
#ifdef DDLO_VISIT_STATS
  if (!g->d_dbg || capacity_points < g->src->n) return fail(DDLO_E_SIZE, "no statistics / capacity too small");
  DDLO_CUDA(cudaStreamSynchronize(g->rt->stream));
  DDLO_CUDA(cudaMemcpy(out, g->d_dbg, (size_t)4 * g->src->n * sizeof(int4), cudaMemcpyDeviceToHost));
  return g->src->n;
#else
  return 0;
#endif
}

// size of the record one align copies back to the host (bench.py's d2h_bytes_per_step)
int ddlo_align_d2h_bytes(void) { return (int)sizeof(AlignOut); }

// ---- host-callable copies of the device math (CPU tests of the exact code the kernels run) --------------
void ddlo_math_sym3_eig(const double* sym6, double* w3, double* V9) {
  Sym3 s{sym6[0], sym6[1], sym6[2], sym6[3], sym6[4], sym6[5]};
  sym3_eig(s, w3, V9);
}
void ddlo_math_regularize(const double* sym6, int method, double* out6) {
  Sym3 s{sym6[0], sym6[1], sym6[2], sym6[3], sym6[4], sym6[5]};
  Sym3 r = regularize_cov(s, method);
  out6[0] = r.xx, out6[1] = r.xy, out6[2] = r.xz, out6[3] = r.yy, out6[4] = r.yz, out6[5] = r.zz;
}
void ddlo_math_ldlt6_solve(const double* A36, const double* rhs6, double* x6) { ldlt6_solve(A36, rhs6, x6); }
void ddlo_math_ldlt6_solve_fast(const double* A36, const double* rhs6, double* x6) { ldlt6_solve_fast(A36, rhs6, x6); }
void ddlo_math_so3_exp(const double* omega3, double* R9) { so3_exp_matrix(omega3, R9); }
void ddlo_math_sym3_inverse(const double* sym6, double* out6) {
  Sym3 s{sym6[0], sym6[1], sym6[2], sym6[3], sym6[4], sym6[5]};
  Sym3 r = sym3_inverse(s);
  out6[0] = r.xx, out6[1] = r.xy, out6[2] = r.xz, out6[3] = r.yy, out6[4] = r.yz, out6[5] = r.zz;
}

}  // extern "C"
