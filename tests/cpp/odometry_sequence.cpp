// BASELINE config C3 with no Python in the frame: the S2S -> S2M odometry loop of OdomNode in C++ on the C ABI,
// keyframes and submaps on the device (ddlo_keyframes_*).
//
// Per frame, the calls of OdomNode (src/odometry/odom.cc):
//   preprocessPoints        :436-478   optional voxel filter of the scan (vf_scan_)
//   initializeInputTarget   :480-516   first frame: S2S target + first keyframe
//   setInputSources         :518-532   S2S source (index), S2M shares cloud + index
//   scanMatching            :745-793   S2S align, propagate, covariance hand-over, swap, submap, S2M align, residuals
//   updateKeyframes         :1067-1150 new-keyframe decision, keyframe cloud (pose applied, optional voxel filter) + covariances
//   getSubmapKeyframes      :1215-1315 k nearest + convex / concave hull keyframes, concatenated on the device
//
//   odometry_sequence scans.bin k thresh_dist thresh_rot_deg knn kcv kcc [voxel_scan voxel_submap [warm_frames]]
//   scans.bin: int32 count, then per scan int32 n and n*4 float32 (x y z 1)
// Prints one line per frame: frame, S2S / S2M iterations and converged flags, new keyframe, submap changed, submap points,
// ms of host wall clock, the 16 entries of T (row-major); then a summary line.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/ddlo_gicp.h"

struct Scan {
  std::vector<float> xyzw;
  int n = 0;
};

static std::vector<Scan> load(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    std::perror(path);
    std::exit(2);
  }
  int count = 0;
  if (std::fread(&count, 4, 1, f) != 1) std::exit(2);
  std::vector<Scan> scans(count);
  for (auto& s : scans) {
    if (std::fread(&s.n, 4, 1, f) != 1) std::exit(2);
    s.xyzw.resize(4 * (size_t)s.n);
    if (std::fread(s.xyzw.data(), 4, s.xyzw.size(), f) != s.xyzw.size()) std::exit(2);
  }
  std::fclose(f);
  return scans;
}

#define CHECK(expr)                                                                 \
  do {                                                                              \
    int rc__ = (expr);                                                              \
    if (rc__ != DDLO_OK) {                                                          \
      std::fprintf(stderr, "%s failed (%d): %s\n", #expr, rc__, ddlo_last_error()); \
      std::exit(1);                                                                 \
    }                                                                               \
  } while (0)

// column-major 4x4 float, like Eigen::Matrix4f
struct M4 {
  float v[16];
  float& operator()(int r, int c) { return v[4 * c + r]; }
  float operator()(int r, int c) const { return v[4 * c + r]; }
};
static M4 identity() {
  M4 m;
  for (int i = 0; i < 16; ++i) m.v[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  return m;
}
static M4 mul(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}
// rotq_ = Eigen::Quaternionf(rotation block of T_), as (w, x, y, z)
static void quaternion(const M4& T, float q[4]) {
  const double t = (double)T(0, 0) + T(1, 1) + T(2, 2);
  if (t > 0) {
    const double s = std::sqrt(t + 1.0) * 2;
    q[0] = (float)(0.25 * s), q[1] = (float)((T(2, 1) - T(1, 2)) / s), q[2] = (float)((T(0, 2) - T(2, 0)) / s), q[3] = (float)((T(1, 0) - T(0, 1)) / s);
  } else {
    int i = 0;
    if (T(1, 1) > T(0, 0)) i = 1;
    if (T(2, 2) > T(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    const double s = std::sqrt(1.0 + T(i, i) - T(j, j) - T(k, k)) * 2;
    q[0] = (float)((T(k, j) - T(j, k)) / s);
    q[1 + i] = (float)(0.25 * s);
    q[1 + j] = (float)((T(j, i) + T(i, j)) / s);
    q[1 + k] = (float)((T(k, i) + T(i, k)) / s);
  }
}

struct Loop {
  ddlo_runtime* rt = nullptr;
  ddlo_gicp *s2s = nullptr, *s2m = nullptr;
  ddlo_keyframes* kf = nullptr;
  M4 T = identity(), T_s2s_prev = identity();
  float thresh_dist = 1.0f, thresh_rot = 15.0f, voxel_scan = 0.0f, voxel_submap = 0.0f;
  int knn = 10, kcv = 10, kcc = 10;
  bool initialised = false;
  int submap_points = 0;
  double* pinned_residuals = nullptr;  // page-locked: lets the residual copy ride on the align's synchronisation
  int pinned_capacity = 0;

  // the keyframe cloud of the current scan (world frame, optional voxel filter), its covariances through the S2S source slot
  void add_keyframe(ddlo_cloud* scan) {
    ddlo_cloud* world = nullptr;
    CHECK(ddlo_cloud_transform(scan, T.v, &world));
    if (voxel_submap > 0.0f) {
      ddlo_cloud* filtered = nullptr;
      CHECK(ddlo_cloud_voxel_filter(world, voxel_submap, voxel_submap, voxel_submap, &filtered));
      ddlo_cloud_release(world);
      world = filtered;
    }
    CHECK(ddlo_gicp_set_input_source(s2s, world, 1));
    CHECK(ddlo_gicp_calculate_source_covariances(s2s));
    ddlo_covs* covs = nullptr;
    CHECK(ddlo_gicp_get_source_covariances(s2s, &covs));
    const float pos[3] = {T(0, 3), T(1, 3), T(2, 3)};
    float q[4];
    quaternion(T, q);
    CHECK(ddlo_keyframes_add(kf, pos, q, world, covs));
    ddlo_covs_release(covs);
    ddlo_cloud_release(world);
  }

  // returns false for the very first frame (no registration)
  bool step(const Scan& scan, ddlo_align_result& r1, ddlo_align_result& r2, int& new_kf, int& changed, std::vector<double>& residuals) {
    ddlo_cloud* cur = nullptr;
    CHECK(ddlo_cloud_create(rt, scan.xyzw.data(), scan.n, 16, &cur));
    if (voxel_scan > 0.0f) {
      ddlo_cloud* filtered = nullptr;
      CHECK(ddlo_cloud_voxel_filter(cur, voxel_scan, voxel_scan, voxel_scan, &filtered));
      ddlo_cloud_release(cur);
      cur = filtered;
    }
    if (!initialised) {
      CHECK(ddlo_gicp_set_input_target(s2s, cur));
      CHECK(ddlo_gicp_calculate_target_covariances(s2s));
      add_keyframe(cur);
      initialised = true;
      ddlo_cloud_release(cur);
      return false;
    }
    // setInputSources
    CHECK(ddlo_gicp_set_input_source(s2s, cur, 1));
    CHECK(ddlo_gicp_set_input_source(s2m, cur, 0));  // registerInputSource; the index travels with the cloud handle
    CHECK(ddlo_gicp_set_source_covariances(s2m, nullptr));
    // scanMatching
    CHECK(ddlo_gicp_align(s2s, nullptr, &r1));
    M4 T_S2S;
    std::memcpy(T_S2S.v, r1.final_transformation, sizeof(T_S2S.v));
    const M4 T_s2s = mul(T_s2s_prev, T_S2S);  // propagateS2S
    ddlo_covs* sc = nullptr;
    CHECK(ddlo_gicp_get_source_covariances(s2s, &sc));
    CHECK(ddlo_gicp_set_source_covariances(s2m, sc));  // gicp_s2m_.source_covs_ = gicp_s2s_.source_covs_
    ddlo_covs_release(sc);
    CHECK(ddlo_gicp_swap_source_and_target(s2s));
    const float cur_pos[3] = {T_s2s(0, 3), T_s2s(1, 3), T_s2s(2, 3)};
    ddlo_cloud* submap = nullptr;
    ddlo_covs* submap_covs = nullptr;
    CHECK(ddlo_keyframes_get_submap(kf, cur_pos, knn, kcv, kcc, (double)thresh_dist, &changed, &submap, &submap_covs, nullptr, 0, nullptr));
    if (changed) {
      CHECK(ddlo_gicp_set_input_target(s2m, submap));
      CHECK(ddlo_gicp_set_target_covariances(s2m, submap_covs));
      CHECK(ddlo_cloud_size(submap, &submap_points));
      ddlo_cloud_release(submap);
      ddlo_covs_release(submap_covs);
    }
    // the S2M align and getResiduals (odom.cc:787-793) behind ONE host synchronisation: both are enqueued, align_finish waits
    int n_cur = 0;
    CHECK(ddlo_cloud_size(cur, &n_cur));
    if (n_cur > pinned_capacity) {
      if (pinned_residuals) CHECK(ddlo_host_free(pinned_residuals));
      pinned_capacity = n_cur + n_cur / 4;
      CHECK(ddlo_host_alloc(sizeof(double) * (size_t)pinned_capacity, reinterpret_cast<void**>(&pinned_residuals)));
    }
    CHECK(ddlo_gicp_align_async(s2m, T_s2s.v));
    CHECK(ddlo_gicp_get_residuals_async(s2m, pinned_residuals, n_cur));
    CHECK(ddlo_gicp_align_finish(s2m, &r2));
    std::memcpy(T.v, r2.final_transformation, sizeof(T.v));
    residuals.assign(pinned_residuals, pinned_residuals + n_cur);
    T_s2s_prev = T;
    // updateKeyframes
    const float pos[3] = {T(0, 3), T(1, 3), T(2, 3)};
    float q[4];
    quaternion(T, q);
    CHECK(ddlo_keyframes_is_new(kf, pos, q, thresh_dist, thresh_rot, &new_kf, nullptr, nullptr, nullptr));
    if (new_kf) add_keyframe(cur);  // enqueued; the next frame's work is ordered behind it on the same stream
    ddlo_cloud_release(cur);
    return true;
  }
};

static Loop make_loop(ddlo_runtime* rt, int k, char** argv, int argc) {
  Loop L;
  L.rt = rt;
  L.thresh_dist = (float)std::atof(argv[3]);
  L.thresh_rot = (float)std::atof(argv[4]);
  L.knn = std::atoi(argv[5]);
  L.kcv = std::atoi(argv[6]);
  L.kcc = std::atoi(argv[7]);
  if (argc > 9) L.voxel_scan = (float)std::atof(argv[8]), L.voxel_submap = (float)std::atof(argv[9]);
  CHECK(ddlo_gicp_create(rt, &L.s2s));
  CHECK(ddlo_gicp_create(rt, &L.s2m));
  for (ddlo_gicp* g : {L.s2s, L.s2m}) {
    ddlo_params p;
    CHECK(ddlo_gicp_get_params(g, &p));
    p.k_correspondences = k;
    CHECK(ddlo_gicp_set_params(g, &p));
  }
  CHECK(ddlo_keyframes_create(rt, &L.kf));
  return L;
}
static void destroy_loop(Loop& L) {
  CHECK(ddlo_runtime_synchronize(L.rt));
  if (L.pinned_residuals) CHECK(ddlo_host_free(L.pinned_residuals));
  CHECK(ddlo_keyframes_destroy(L.kf));
  CHECK(ddlo_gicp_destroy(L.s2s));
  CHECK(ddlo_gicp_destroy(L.s2m));
}

int main(int argc, char** argv) {
  if (argc < 8) {
    std::fprintf(stderr, "usage: %s scans.bin k thresh_dist thresh_rot_deg knn kcv kcc [voxel_scan voxel_submap [warm_frames]]\n", argv[0]);
    return 2;
  }
  const std::vector<Scan> scans = load(argv[1]);
  const int k = std::atoi(argv[2]);
  const int warm = argc > 10 ? std::atoi(argv[10]) : 0;
  ddlo_runtime* rt = nullptr;
  CHECK(ddlo_runtime_create(0, &rt));
  if (warm > 0) {  // a short untimed run first: allocator pools, first launches
    Loop W = make_loop(rt, k, argv, argc);
    ddlo_align_result a, b;
    int nk = 0, ch = 0;
    std::vector<double> res;
    for (int f = 0; f < std::min<int>(warm, (int)scans.size()); ++f) W.step(scans[f], a, b, nk, ch, res);
    destroy_loop(W);
  }
  Loop L = make_loop(rt, k, argv, argc);
  std::vector<double> ms;
  int keyframes = 0;
  for (size_t f = 0; f < scans.size(); ++f) {
    ddlo_align_result r1, r2;
    int new_kf = 0, changed = 0;
    std::vector<double> residuals;
    const auto t0 = std::chrono::steady_clock::now();
    const bool registered = L.step(scans[f], r1, r2, new_kf, changed, residuals);
    const double dt = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!registered) continue;
    ms.push_back(dt);
    double rsum = 0.0;
    for (double v : residuals) rsum += v;
    std::printf("frame %zu %d %d %d %d %d %d %d %.4f %.12g", f, r1.nr_iterations, r2.nr_iterations, (r1.flags & DDLO_FLAG_CONVERGED) ? 1 : 0,
                (r2.flags & DDLO_FLAG_CONVERGED) ? 1 : 0, new_kf, changed, L.submap_points, dt, residuals.empty() ? 0.0 : rsum / (double)residuals.size());
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) std::printf(" %.9g", L.T(r, c));
    std::printf("\n");
  }
  CHECK(ddlo_keyframes_count(L.kf, &keyframes));
  std::vector<double> sorted = ms;
  std::sort(sorted.begin(), sorted.end());
  double mean = 0.0;
  for (double v : ms) mean += v;
  mean /= std::max<size_t>(ms.size(), 1);
  auto pct = [&](double p) { return sorted.empty() ? 0.0 : sorted[std::min(sorted.size() - 1, (size_t)(p * (sorted.size() - 1) + 0.5))]; };
  std::printf("summary frames %zu keyframes %d mean_ms %.4f p50_ms %.4f p99_ms %.4f max_ms %.4f\n", ms.size(), keyframes, mean, pct(0.5), pct(0.99),
              sorted.empty() ? 0.0 : sorted.back());
  destroy_loop(L);
  CHECK(ddlo_runtime_destroy(rt));
  return 0;
}
