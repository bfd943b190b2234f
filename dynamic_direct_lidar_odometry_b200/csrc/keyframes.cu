// Device-resident keyframe store and submap builder, host logic in C++.
//
// Replaces, around the registration path, what OdomNode keeps in host vectors:
//   keyframes_ / keyframe_normals_                    odom.h:105-108        -> clouds and covariance vectors stay in HBM
//   updateKeyframes' new-keyframe decision            odom.cc:1067-1126     -> ddlo_keyframes_is_new
//   keyframes_.push_back + keyframe_normals_.push_back odom.cc:502-510, 1139-1149 -> ddlo_keyframes_add
//   getSubmapKeyframes                                odom.cc:1215-1315     -> ddlo_keyframes_get_submap
//     pushSubmapIndices (k nearest, ties included)    odom.cc:1178-1213
//     computeConvexHull / computeConcaveHull          odom.cc:993-1065      -> hull.hpp (host; a few hundred poses)
//     `*submap_cloud += *keyframes_[k].second` and the submap_normals_ inserts (:1298-1313) -> device-side concatenation
// The selection is a handful of scalar operations per keyframe and runs on the host; points and covariances never
// leave the device: a new submap costs two device-to-device concatenations instead of the reference's host
// concatenation, kd-tree rebuild and 64 MB covariance copy.
#include <algorithm>
#include <cmath>
#include <queue>
#include <vector>

#include "engine.cuh"
#include "hull.hpp"

using namespace ddlo;

struct ddlo_keyframes {
  ddlo_runtime* rt = nullptr;
  struct Keyframe {
    float pos[3];
    float q[4];  // w, x, y, z
    ddlo_cloud* cloud = nullptr;
    ddlo_covs* covs = nullptr;
  };
  std::vector<Keyframe> kf;
  std::vector<int> submap_prev;  // submap_kf_idx_prev_
  std::vector<int> convex, concave;  // keyframe_convex_ / keyframe_concave_ (kept between calls like the reference's members)
  int concave_dimension = 2;
};

// OdomNode::pushSubmapIndices (odom.cc:1178-1213): every frame whose distance is <= the k-th smallest distance
static void push_submap_indices(const std::vector<float>& dists, int k, const std::vector<int>& frames, std::vector<int>& out) {
  if (dists.empty() || k <= 0) return;  // (the reference never asks for k = 0: it would read the top of an empty heap)
  std::priority_queue<float> pq;
  for (float d : dists) {
    if ((int)pq.size() >= k && pq.top() > d) {
      pq.push(d);
      pq.pop();
    } else if ((int)pq.size() < k) {
      pq.push(d);
    }
  }
  const float kth = pq.top();
  for (size_t i = 0; i < dists.size(); ++i)
    if (dists[i] <= kth) out.push_back(frames[i]);
}

// float arithmetic of the reference: sqrt(pow(a - b, 2) + ...) with float operands is evaluated in double (std::pow
// of a float and an int promotes to double) and rounded to float on assignment
static float pose_distance(const float* a, const float* b) {
  const double dx = (double)(a[0] - b[0]), dy = (double)(a[1] - b[1]), dz = (double)(a[2] - b[2]);
  return (float)std::sqrt(dx * dx + dy * dy + dz * dz);
}

extern "C" {

int ddlo_keyframes_create(ddlo_runtime* rt, ddlo_keyframes** out) {
  if (!rt || !out) return fail(DDLO_E_INVALID, "null argument");
  ddlo_keyframes* k = new (std::nothrow) ddlo_keyframes();
  if (!k) return fail(DDLO_E_INVALID, "out of host memory");
  k->rt = rt;
  *out = k;
  return DDLO_OK;
}

int ddlo_keyframes_destroy(ddlo_keyframes* k) {
  if (!k) return DDLO_OK;
  for (auto& f : k->kf) {
    ddlo_cloud_release(f.cloud);
    ddlo_covs_release(f.covs);
  }
  delete k;
  return DDLO_OK;
}

int ddlo_keyframes_count(const ddlo_keyframes* k, int* n) {
  if (!k || !n) return fail(DDLO_E_INVALID, "null argument");
  *n = (int)k->kf.size();
  return DDLO_OK;
}

int ddlo_keyframes_add(ddlo_keyframes* k, const float* position_xyz, const float* rotation_wxyz, ddlo_cloud* cloud_world, ddlo_covs* covs) {
  if (!k || !position_xyz || !rotation_wxyz || !cloud_world || !covs) return fail(DDLO_E_INVALID, "null argument");
  if (cloud_world->rt != k->rt || covs->rt != k->rt) return fail(DDLO_E_INVALID, "keyframe of another runtime");
  if (cloud_world->n != covs->n) return fail(DDLO_E_SIZE, "keyframe cloud and covariances differ in size");
  ddlo_keyframes::Keyframe f;
  for (int i = 0; i < 3; ++i) f.pos[i] = position_xyz[i];
  for (int i = 0; i < 4; ++i) f.q[i] = rotation_wxyz[i];
  f.cloud = cloud_world;
  f.covs = covs;
  ddlo_cloud_retain(cloud_world);
  ddlo_covs_retain(covs);
  k->kf.push_back(f);
  return DDLO_OK;
}

int ddlo_keyframes_get(ddlo_keyframes* k, int index, float* position_xyz, float* rotation_wxyz, ddlo_cloud** cloud, ddlo_covs** covs) {
  if (!k) return fail(DDLO_E_INVALID, "null argument");
  if (index < 0 || index >= (int)k->kf.size()) return fail(DDLO_E_INVALID, "no such keyframe");
  const auto& f = k->kf[index];
  if (position_xyz)
    for (int i = 0; i < 3; ++i) position_xyz[i] = f.pos[i];
  if (rotation_wxyz)
    for (int i = 0; i < 4; ++i) rotation_wxyz[i] = f.q[i];
  if (cloud) {
    *cloud = f.cloud;
    ddlo_cloud_retain(f.cloud);
  }
  if (covs) {
    *covs = f.covs;
    ddlo_covs_retain(f.covs);
  }
  return DDLO_OK;
}

// updateKeyframes (odom.cc:1067-1126) up to the decision
int ddlo_keyframes_is_new(const ddlo_keyframes* k, const float* position_xyz, const float* rotation_wxyz, float thresh_dist, float thresh_rot_deg,
                          int* is_new, int* closest_index, float* closest_distance, float* rotation_deg) {
  if (!k || !position_xyz || !rotation_wxyz || !is_new) return fail(DDLO_E_INVALID, "null argument");
  if (k->kf.empty()) return fail(DDLO_E_NOT_READY, "no keyframes yet");
  float closest_d = std::numeric_limits<float>::infinity();
  int closest_idx = 0, num_nearby = 0;
  for (size_t i = 0; i < k->kf.size(); ++i) {
    const float d = pose_distance(position_xyz, k->kf[i].pos);
    if ((double)d <= (double)thresh_dist * 1.5) ++num_nearby;
    if (d < closest_d) closest_d = d, closest_idx = (int)i;
  }
  const float dd = pose_distance(position_xyz, k->kf[closest_idx].pos);
  // dq = rotq_ * closest_pose_r.inverse()  (Eigen::Quaternionf: inverse = conjugate / squared norm)
  const float* a = rotation_wxyz;
  const float* c = k->kf[closest_idx].q;
  const float n2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3];
  float b[4] = {c[0] / n2, -c[1] / n2, -c[2] / n2, -c[3] / n2};
  if (!(n2 > 0.0f)) b[0] = b[1] = b[2] = b[3] = 0.0f;  // Eigen returns the zero quaternion
  const float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  const float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  const float y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  const float z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
  const float theta_rad = (float)(2. * std::atan2(std::sqrt((double)x * x + (double)y * y + (double)z * z), (double)w));
  const float theta_deg = (float)((double)theta_rad * (180.0 / M_PI));
  bool nk = false;
  if (std::fabs(dd) > thresh_dist || std::fabs(theta_deg) > thresh_rot_deg) nk = true;
  if (std::fabs(dd) <= thresh_dist) nk = false;
  if (std::fabs(dd) <= thresh_dist && std::fabs(theta_deg) > thresh_rot_deg && num_nearby <= 1) nk = true;
  *is_new = nk ? 1 : 0;
  if (closest_index) *closest_index = closest_idx;
  if (closest_distance) *closest_distance = dd;
  if (rotation_deg) *rotation_deg = theta_deg;
  return DDLO_OK;
}

int ddlo_keyframes_get_submap(ddlo_keyframes* k, const float* current_xyz, int submap_knn, int submap_kcv, int submap_kcc, double concave_alpha,
                              int* changed, ddlo_cloud** cloud, ddlo_covs** covs, int* indices, int capacity, int* n_indices) {
  if (!k || !current_xyz || !changed) return fail(DDLO_E_INVALID, "null argument");
  if (k->kf.empty()) return fail(DDLO_E_NOT_READY, "no keyframes yet");
  if (cloud) *cloud = nullptr;
  if (covs) *covs = nullptr;
  const int n = (int)k->kf.size();
  std::vector<int> sel;
  std::vector<float> ds(n);
  std::vector<int> all(n);
  for (int i = 0; i < n; ++i) ds[i] = pose_distance(current_xyz, k->kf[i].pos), all[i] = i;
  push_submap_indices(ds, submap_knn, all, sel);
  std::vector<hull::P3> pts(n);
  for (int i = 0; i < n; ++i) pts[i] = {(double)k->kf[i].pos[0], (double)k->kf[i].pos[1], (double)k->kf[i].pos[2]};
  // computeConvexHull: "at least 4 keyframes"; below that keyframe_convex_ keeps its previous content (empty)
  if (n >= 4) k->convex = hull::convex_hull_indices(pts);
  std::vector<float> cds;
  for (int c : k->convex) cds.push_back(ds[c]);
  push_submap_indices(cds, submap_kcv, k->convex, sel);
  // computeConcaveHull: "at least 5 keyframes"
  if (n >= 5) k->concave = hull::concave_hull_indices(pts, concave_alpha, &k->concave_dimension);
  std::vector<float> ccds;
  for (int c : k->concave) ccds.push_back(ds[c]);
  push_submap_indices(ccds, submap_kcc, k->concave, sel);
  std::sort(sel.begin(), sel.end());
  sel.erase(std::unique(sel.begin(), sel.end()), sel.end());
  if (n_indices) *n_indices = (int)sel.size();
  if (indices)
    for (int i = 0; i < std::min<int>((int)sel.size(), capacity); ++i) indices[i] = sel[i];
  if (sel == k->submap_prev) {
    *changed = 0;  // submap_hasChanged_ = false: the caller keeps its target
    return DDLO_OK;
  }
  *changed = 1;
  if (cloud && covs) {
    std::vector<ddlo_cloud*> cs;
    std::vector<ddlo_covs*> vs;
    for (int i : sel) cs.push_back(k->kf[i].cloud), vs.push_back(k->kf[i].covs);
    DDLO_TRY(ddlo_cloud_concat(k->rt, cs.data(), (int)cs.size(), cloud));
    const int rc = ddlo_covs_concat(k->rt, vs.data(), (int)vs.size(), covs);
    if (rc != DDLO_OK) {
      ddlo_cloud_release(*cloud);
      *cloud = nullptr;
      return rc;
    }
  }
  k->submap_prev = sel;
  return DDLO_OK;
}

// the hull vertex sets of the last ddlo_keyframes_get_submap (keyframe_convex_, keyframe_concave_), and the dimension
// the concave hull was taken in (3: not computed, see hull.hpp)
int ddlo_keyframes_hulls(const ddlo_keyframes* k, int* convex, int* n_convex, int* concave, int* n_concave, int capacity, int* concave_dimension) {
  if (!k) return fail(DDLO_E_INVALID, "null argument");
  if (n_convex) *n_convex = (int)k->convex.size();
  if (n_concave) *n_concave = (int)k->concave.size();
  if (convex)
    for (int i = 0; i < std::min<int>((int)k->convex.size(), capacity); ++i) convex[i] = k->convex[i];
  if (concave)
    for (int i = 0; i < std::min<int>((int)k->concave.size(), capacity); ++i) concave[i] = k->concave[i];
  if (concave_dimension) *concave_dimension = k->concave_dimension;
  return DDLO_OK;
}

// host-callable hull functions (CPU tests compare them with Qhull through scipy): xyz = n * 3 doubles;
// return the number of hull vertices (indices written up to capacity), concave: -3 when the points are 3-dimensional
int ddlo_hull_convex(const double* xyz, int n, int* out_indices, int capacity) {
  std::vector<hull::P3> p(std::max(n, 0));
  for (int i = 0; i < n; ++i) p[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  const std::vector<int> h = hull::convex_hull_indices(p);
  for (int i = 0; i < std::min<int>((int)h.size(), capacity); ++i) out_indices[i] = h[i];
  return (int)h.size();
}
int ddlo_hull_concave(const double* xyz, int n, double alpha, int* out_indices, int capacity) {
  std::vector<hull::P3> p(std::max(n, 0));
  for (int i = 0; i < n; ++i) p[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  int dim = 2;
  const std::vector<int> h = hull::concave_hull_indices(p, alpha, &dim);
  if (dim == 3) return -3;
  for (int i = 0; i < std::min<int>((int)h.size(), capacity); ++i) out_indices[i] = h[i];
  return (int)h.size();
}
int ddlo_hull_dimension(const double* xyz, int n) {
  std::vector<hull::P3> p(std::max(n, 0));
  for (int i = 0; i < n; ++i) p[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  return n >= 3 ? hull::input_dimension(p) : 2;
}

}  // extern "C"
