// nano_gicp::NanoGICP on the B200 — header-only C++17 shim over the C ABI (include/ddlo_gicp.h).
//
// Drop-in for the reference's registration engine
//   /root/reference/dynamic_direct_lidar_odometry/include/nano_gicp/nano_gicp.hpp:58-148
//   /root/reference/dynamic_direct_lidar_odometry/include/nano_gicp/lsq_registration.hpp:60-128
//   /root/reference/dynamic_direct_lidar_odometry/include/nano_gicp/nanoflann.hpp:53-203
// Same class, method and member names, same state rules (setInputSource clears the source
// covariances, swapSourceAndTarget moves clouds + trees + covariances, the public members
// source_kdtree_ / target_kdtree_ / source_covs_ / target_covs_ can be shared between engines the
// way OdomNode does, odom.cc:527-531,765), so OdomNode compiles against it unchanged.  All
// arithmetic runs in libddlo_gicp_b200.so on the GPU; this header only keeps handles in step.
//
// Types.  With PCL and Eigen on the include path the engine takes pcl::PointCloud<PointT>::ConstPtr
// and Eigen matrices like the reference.  Without them (as in this repository's build image, which
// has neither) it falls back to the minimal stand-ins defined below with the same memory layout
// (32-byte XYZI points, column-major 4x4), which is what the repository's own C++ test uses.  The
// PCL/Eigen branch could not be compiled here and is kept deliberately thin.
#pragma once

#include <array>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ddlo_gicp.h"

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && !defined(DDLO_NO_EIGEN)
#define DDLO_HAVE_EIGEN 1
#include <Eigen/Core>
#include <Eigen/StdVector>
#endif
#if __has_include(<pcl/point_cloud.h>) && !defined(DDLO_NO_PCL)
#define DDLO_HAVE_PCL 1
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#endif
#endif

namespace ddlo_shim {

#ifdef DDLO_HAVE_EIGEN
using Matrix4f = Eigen::Matrix4f;
using Matrix4d = Eigen::Matrix4d;
using Matrix6d = Eigen::Matrix<double, 6, 6>;
using Vector3f = Eigen::Vector3f;
using Matrix4dVector = std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>;
inline float* data(Matrix4f& m) { return m.data(); }
inline const float* data(const Matrix4f& m) { return m.data(); }
inline double* data(Matrix6d& m) { return m.data(); }
inline Matrix4f identity4f() { return Matrix4f::Identity(); }
#else
// column-major fixed matrices with Eigen's element access, just enough for the engine's interface
template <class T, int N>
struct Mat {
  T v[N * N];
  T& operator()(int r, int c) { return v[c * N + r]; }
  const T& operator()(int r, int c) const { return v[c * N + r]; }
  T* data() { return v; }
  const T* data() const { return v; }
  static Mat Identity() {
    Mat m;
    for (int i = 0; i < N * N; ++i) m.v[i] = (i % (N + 1) == 0) ? T(1) : T(0);
    return m;
  }
  static Mat Zero() {
    Mat m;
    for (int i = 0; i < N * N; ++i) m.v[i] = T(0);
    return m;
  }
};
using Matrix4f = Mat<float, 4>;
using Matrix4d = Mat<double, 4>;
using Matrix6d = Mat<double, 6>;
struct Vector3f {
  float v[3];
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
};
using Matrix4dVector = std::vector<Matrix4d>;
inline float* data(Matrix4f& m) { return m.data(); }
inline const float* data(const Matrix4f& m) { return m.data(); }
inline double* data(Matrix6d& m) { return m.data(); }
inline Matrix4f identity4f() { return Matrix4f::Identity(); }
#endif

#ifdef DDLO_HAVE_PCL
template <class PointT>
using Cloud = pcl::PointCloud<PointT>;
using PointXYZI = pcl::PointXYZI;
#else
// pcl::PointXYZI: x, y, z, 1.0f | intensity + padding = 32 bytes
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, w = 1.0f;
  float intensity = 0, pad[3] = {0, 0, 0};
};
template <class PointT>
struct Cloud {
  using Ptr = std::shared_ptr<Cloud<PointT>>;
  using ConstPtr = std::shared_ptr<const Cloud<PointT>>;
  std::vector<PointT> points;
  size_t size() const { return points.size(); }
  void resize(size_t n) { points.resize(n); }
  const PointT& at(size_t i) const { return points.at(i); }
  PointT& at(size_t i) { return points.at(i); }
};
#endif

inline void check(int rc, const char* what) {
  if (rc != DDLO_OK) throw std::runtime_error(std::string(what) + ": " + ddlo_last_error());
}

// one runtime (device + stream) per process and device, shared by every engine like the single
// address space the reference lives in
inline ddlo_runtime* runtime(int device = 0) {
  struct Holder {
    ddlo_runtime* rt[16] = {};
    ~Holder() {
      for (auto* r : rt)
        if (r) ddlo_runtime_destroy(r);
    }
  };
  static Holder h;
  if (device < 0 || device >= 16) throw std::runtime_error("ddlo: device index out of range");
  if (!h.rt[device]) check(ddlo_runtime_create(device, &h.rt[device]), "ddlo_runtime_create");
  return h.rt[device];
}

// shared ownership of the opaque C handles
struct CloudHandle {
  ddlo_cloud* h = nullptr;
  explicit CloudHandle(ddlo_cloud* c) : h(c) {}
  ~CloudHandle() { ddlo_cloud_release(h); }
  CloudHandle(const CloudHandle&) = delete;
  CloudHandle& operator=(const CloudHandle&) = delete;
};
struct CovsHandle {
  ddlo_covs* h = nullptr;
  explicit CovsHandle(ddlo_covs* c) : h(c) {}
  ~CovsHandle() { ddlo_covs_release(h); }
  CovsHandle(const CovsHandle&) = delete;
  CovsHandle& operator=(const CovsHandle&) = delete;
};

}  // namespace ddlo_shim

namespace nanoflann {

// nanoflann::KdTreeFLANN<PointT> (nanoflann.hpp:53-203): the index over one cloud.  On the device the
// index is a property of the uploaded cloud, so this object owns the device copy of the cloud.
template <class PointT>
class KdTreeFLANN {
 public:
  using PointCloudConstPtr = typename ddlo_shim::Cloud<PointT>::ConstPtr;

  void setInputCloud(const PointCloudConstPtr& cloud) {
    host_ = cloud;
    device_.reset();
    if (!cloud || cloud->size() == 0) return;
    ddlo_cloud* c = nullptr;
    ddlo_shim::check(ddlo_cloud_create(ddlo_shim::runtime(), reinterpret_cast<const float*>(cloud->points.data()), (int)cloud->size(),
                                       (int)sizeof(PointT), &c),
                     "ddlo_cloud_create");
    device_ = std::make_shared<ddlo_shim::CloudHandle>(c);
    ddlo_shim::check(ddlo_cloud_build_index(c), "ddlo_cloud_build_index");
  }
  PointCloudConstPtr getInputCloud() const { return host_; }

  int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    k_indices.assign(k, -1);
    k_sqr_distances.assign(k, std::numeric_limits<float>::infinity());
    if (!device_) return 0;
    int count = 0;
    ddlo_shim::check(ddlo_cloud_knn(device_->h, &point.x, 1, (int)sizeof(PointT), k, k_indices.data(), k_sqr_distances.data(), &count),
                     "ddlo_cloud_knn");
    return count;
  }

  // shim internals
  const std::shared_ptr<ddlo_shim::CloudHandle>& device() const { return device_; }

 private:
  PointCloudConstPtr host_;
  std::shared_ptr<ddlo_shim::CloudHandle> device_;
};

}  // namespace nanoflann

namespace nano_gicp {

// gicp/gicp_settings.hpp:47-54
enum class RegularizationMethod { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS };
// lsq_registration.hpp:54-58
enum class LSQ_OPTIMIZER_TYPE { GaussNewton, LevenbergMarquardt };

// std::vector<Eigen::Matrix4d> that lives on the device.  size / clear / copy-assignment (shares the
// buffer: `s2m.source_covs_ = s2s.source_covs_`, odom.cc:765, is free) / assignment from and
// conversion to the host vector type (keyframe_normals_, submap_normals_).
class DeviceCovariances {
 public:
  DeviceCovariances() = default;
  DeviceCovariances(const ddlo_shim::Matrix4dVector& v) { *this = v; }
  DeviceCovariances& operator=(const ddlo_shim::Matrix4dVector& v) {
    h_.reset();
    if (!v.empty()) {
      ddlo_covs* c = nullptr;
      ddlo_shim::check(ddlo_covs_from_host(ddlo_shim::runtime(), reinterpret_cast<const double*>(v.data()), (int)v.size(), &c),
                       "ddlo_covs_from_host");
      h_ = std::make_shared<ddlo_shim::CovsHandle>(c);
    }
    return *this;
  }
  size_t size() const {
    int n = 0;
    if (h_) ddlo_covs_size(h_->h, &n);
    return (size_t)n;
  }
  bool empty() const { return size() == 0; }
  void clear() { h_.reset(); }
  void swap(DeviceCovariances& o) { h_.swap(o.h_); }
  operator ddlo_shim::Matrix4dVector() const {
    ddlo_shim::Matrix4dVector v(size());
    if (h_) ddlo_shim::check(ddlo_covs_to_host(h_->h, reinterpret_cast<double*>(v.data())), "ddlo_covs_to_host");
    return v;
  }
  // concatenation on the device (submap assembly, odom.cc:1298-1313)
  static DeviceCovariances concat(const std::vector<DeviceCovariances>& parts) {
    std::vector<ddlo_covs*> hs;
    for (const auto& p : parts)
      if (p.h_) hs.push_back(p.h_->h);
    DeviceCovariances out;
    if (hs.empty()) return out;
    ddlo_covs* c = nullptr;
    ddlo_shim::check(ddlo_covs_concat(ddlo_shim::runtime(), hs.data(), (int)hs.size(), &c), "ddlo_covs_concat");
    out.h_ = std::make_shared<ddlo_shim::CovsHandle>(c);
    return out;
  }
  ddlo_covs* handle() const { return h_ ? h_->h : nullptr; }
  void adopt(ddlo_covs* c) { h_ = c ? std::make_shared<ddlo_shim::CovsHandle>(c) : nullptr; }

 private:
  std::shared_ptr<ddlo_shim::CovsHandle> h_;
};

template <typename PointSource, typename PointTarget>
class NanoGICP {
 public:
  using Scalar = float;
  using Matrix4 = ddlo_shim::Matrix4f;
  using PointCloudSource = ddlo_shim::Cloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = ddlo_shim::Cloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;

  NanoGICP() {
    ddlo_shim::check(ddlo_gicp_create(ddlo_shim::runtime(), &g_), "ddlo_gicp_create");
    ddlo_params_default(&p_);
    source_kdtree_.reset(new nanoflann::KdTreeFLANN<PointSource>);
    target_kdtree_.reset(new nanoflann::KdTreeFLANN<PointTarget>);
    final_transformation_ = ddlo_shim::identity4f();
    final_hessian_ = ddlo_shim::Matrix6d::Identity();
  }
  virtual ~NanoGICP() { ddlo_gicp_destroy(g_); }
  NanoGICP(const NanoGICP&) = delete;
  NanoGICP& operator=(const NanoGICP&) = delete;

  // the C-ABI engine behind this object (for the calls that take an engine, e.g. ddlo_gicp_segment_scan)
  ddlo_gicp* handle() const { return g_; }

  // ---- knobs (nano_gicp.hpp:83-85, lsq_registration.hpp:89-91, pcl::Registration) -------------
  void setNumThreads(int) {}  // OpenMP threads: no meaning on the device
  void setCorrespondenceRandomness(int k) { p_.k_correspondences = k; }
  void setRegularizationMethod(RegularizationMethod m) { p_.regularization_method = (int)m; }
  void setMaxCorrespondenceDistance(double d) { p_.max_correspondence_distance = d; }
  void setMaximumIterations(int n) { p_.max_iterations = n; }
  void setTransformationEpsilon(double e) { p_.transformation_epsilon = e; }
  void setRotationEpsilon(double e) { p_.rotation_epsilon = e; }
  void setInitialLambdaFactor(double f) { p_.lm_init_lambda_factor = f; }
  void setDebugPrint(bool) {}
  // set by OdomNode, ignored by nano_gicp (odom.cc:96-98,104-112)
  void setEuclideanFitnessEpsilon(double) {}
  void setRANSACIterations(int) {}
  void setRANSACOutlierRejectionThreshold(double) {}
  template <class TreePtr>
  void setSearchMethodSource(const TreePtr&, bool = false) {}
  template <class TreePtr>
  void setSearchMethodTarget(const TreePtr&, bool = false) {}

  // ---- state plumbing (nano_gicp_impl.hpp:98-181) --------------------------------------------------
  virtual void swapSourceAndTarget() {
    input_.swap(target_);
    source_kdtree_.swap(target_kdtree_);
    source_covs_.swap(target_covs_);
    have_correspondences_ = false;
  }
  virtual void clearSource() {
    input_.reset();
    source_covs_.clear();
  }
  virtual void clearTarget() {
    target_.reset();
    target_covs_.clear();
  }
  virtual void registerInputSource(const PointCloudSourceConstPtr& cloud) {
    if (input_ == cloud) return;
    input_ = cloud;
  }
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) {
    if (input_ == cloud) return;
    input_ = cloud;
    source_kdtree_->setInputCloud(cloud);
    source_covs_.clear();
  }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    if (target_ == cloud) return;
    target_ = cloud;
    target_kdtree_->setInputCloud(cloud);
    target_covs_.clear();
  }
  virtual void setSourceCovariances(const ddlo_shim::Matrix4dVector& covs) { source_covs_ = covs; }
  virtual void setTargetCovariances(const ddlo_shim::Matrix4dVector& covs) { target_covs_ = covs; }
  virtual void setSourceCovariances(const DeviceCovariances& covs) { source_covs_ = covs; }
  virtual void setTargetCovariances(const DeviceCovariances& covs) { target_covs_ = covs; }
  virtual bool calculateSourceCovariances() { return calculate_covariances(input_, *source_kdtree_, source_covs_); }
  virtual bool calculateTargetCovariances() { return calculate_covariances(target_, *target_kdtree_, target_covs_); }
  const DeviceCovariances& getSourceCovariances() const { return source_covs_; }
  const DeviceCovariances& getTargetCovariances() const { return target_covs_; }

  // ---- registration (pcl::Registration::align -> computeTransformation) ----------------------------
  void align(PointCloudSource& output) { align(output, ddlo_shim::identity4f()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    converged_ = false;
    if (!input_ || !target_) return;  // PCL's initCompute fails quietly
    push_state();
    ddlo_align_result r;
    ddlo_shim::check(ddlo_gicp_align(g_, ddlo_shim::data(guess), &r), "ddlo_gicp_align");
    // covariances computed inside align belong to the engine's public members afterwards (:186-193)
    pull_covariances();
    std::memcpy(ddlo_shim::data(final_transformation_), r.final_transformation, sizeof(r.final_transformation));
    std::memcpy(ddlo_shim::data(final_hessian_), r.final_hessian, sizeof(r.final_hessian));
    converged_ = (r.flags & DDLO_FLAG_CONVERGED) != 0;
    nr_iterations_ = r.nr_iterations;
    lm_failed_ = (r.flags & DDLO_FLAG_LM_FAILED) != 0;
    have_correspondences_ = true;
    if (compute_output_) {
      ddlo_cloud* out = nullptr;
      ddlo_shim::check(ddlo_gicp_aligned_cloud(g_, &out), "ddlo_gicp_aligned_cloud");
      ddlo_shim::CloudHandle guard(out);
      std::vector<float> xyzw(4 * input_->size());
      ddlo_shim::check(ddlo_cloud_download(out, xyzw.data()), "ddlo_cloud_download");
      output.points.assign(input_->points.begin(), input_->points.end());
      for (size_t i = 0; i < input_->size(); ++i) {
        output.points[i].x = xyzw[4 * i];
        output.points[i].y = xyzw[4 * i + 1];
        output.points[i].z = xyzw[4 * i + 2];
      }
    }
  }
  // The reference always fills `output` (lsq_registration_impl.hpp:125); OdomNode throws it away
  // (odom.cc:752-790).  Turning it off saves a 1 MB read-back per align.
  void setComputeOutputCloud(bool on) { compute_output_ = on; }

  Matrix4 getFinalTransformation() const { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  const ddlo_shim::Matrix6d& getFinalHessian() const { return final_hessian_; }
  int getNrIterations() const { return nr_iterations_; }
  bool lmFailed() const { return lm_failed_; }

  void getResiduals(std::vector<ddlo_shim::Vector3f>& residuals, const Matrix4& trans) {
    residuals.resize(input_ ? input_->size() : 0);
    if (residuals.empty()) return;
    require_correspondences();
    ddlo_shim::check(ddlo_gicp_get_residual_vectors(g_, ddlo_shim::data(trans), reinterpret_cast<float*>(residuals.data()), (int)residuals.size()),
                     "ddlo_gicp_get_residual_vectors");
  }
  void getResiduals(std::vector<double>& residuals, const Matrix4&) {
    residuals.resize(input_ ? input_->size() : 0);
    if (residuals.empty()) return;
    require_correspondences();
    ddlo_shim::check(ddlo_gicp_get_residuals(g_, residuals.data(), (int)residuals.size()), "ddlo_gicp_get_residuals");
  }

 protected:
  template <class PointT>
  bool calculate_covariances(const typename ddlo_shim::Cloud<PointT>::ConstPtr& cloud, nanoflann::KdTreeFLANN<PointT>& kdtree,
                             DeviceCovariances& covariances) {
    if (!cloud) return false;
    if (kdtree.getInputCloud() != cloud) kdtree.setInputCloud(cloud);
    if (!kdtree.device()) return false;
    ddlo_covs* c = nullptr;
    ddlo_shim::check(ddlo_covs_compute(kdtree.device()->h, p_.k_correspondences, p_.regularization_method, &c), "ddlo_covs_compute");
    covariances.adopt(c);
    return true;
  }

  // the device copy of `cloud`: the one the kd-tree member holds if it is on that cloud, else a plain upload
  template <class PointT>
  std::shared_ptr<ddlo_shim::CloudHandle> device_cloud(const typename ddlo_shim::Cloud<PointT>::ConstPtr& cloud,
                                                       const std::shared_ptr<nanoflann::KdTreeFLANN<PointT>>& tree,
                                                       std::shared_ptr<ddlo_shim::CloudHandle>& cache,
                                                       typename ddlo_shim::Cloud<PointT>::ConstPtr& cache_key) {
    if (tree && tree->getInputCloud() == cloud && tree->device()) return tree->device();
    if (cache && cache_key == cloud) return cache;
    ddlo_cloud* c = nullptr;
    ddlo_shim::check(ddlo_cloud_create(ddlo_shim::runtime(), reinterpret_cast<const float*>(cloud->points.data()), (int)cloud->size(),
                                       (int)sizeof(PointT), &c),
                     "ddlo_cloud_create");
    cache = std::make_shared<ddlo_shim::CloudHandle>(c);
    cache_key = cloud;
    return cache;
  }

  // bring the C engine in line with the (freely assignable) public members
  void push_state() {
    ddlo_shim::check(ddlo_gicp_set_params(g_, &p_), "ddlo_gicp_set_params");
    auto s = device_cloud<PointSource>(input_, source_kdtree_, src_upload_, src_upload_key_);
    auto t = device_cloud<PointTarget>(target_, target_kdtree_, tgt_upload_, tgt_upload_key_);
    ddlo_shim::check(ddlo_gicp_set_input_source(g_, s->h, 0), "ddlo_gicp_set_input_source");
    ddlo_shim::check(ddlo_gicp_set_input_target(g_, t->h), "ddlo_gicp_set_input_target");
    ddlo_shim::check(ddlo_gicp_set_source_covariances(g_, source_covs_.handle()), "ddlo_gicp_set_source_covariances");
    ddlo_shim::check(ddlo_gicp_set_target_covariances(g_, target_covs_.handle()), "ddlo_gicp_set_target_covariances");
  }
  void pull_covariances() {
    ddlo_covs* c = nullptr;
    ddlo_shim::check(ddlo_gicp_get_source_covariances(g_, &c), "ddlo_gicp_get_source_covariances");
    if (c != source_covs_.handle()) source_covs_.adopt(c); else ddlo_covs_release(c);
    c = nullptr;
    ddlo_shim::check(ddlo_gicp_get_target_covariances(g_, &c), "ddlo_gicp_get_target_covariances");
    if (c != target_covs_.handle()) target_covs_.adopt(c); else ddlo_covs_release(c);
  }
  void require_correspondences() const {
    if (!have_correspondences_) throw std::runtime_error("NanoGICP::getResiduals before align");
  }

 public:
  // nano_gicp.hpp:132-136
  std::shared_ptr<nanoflann::KdTreeFLANN<PointSource>> source_kdtree_;
  std::shared_ptr<nanoflann::KdTreeFLANN<PointTarget>> target_kdtree_;
  DeviceCovariances source_covs_;
  DeviceCovariances target_covs_;

 protected:
  ddlo_gicp* g_ = nullptr;
  ddlo_params p_{};
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  std::shared_ptr<ddlo_shim::CloudHandle> src_upload_, tgt_upload_;
  PointCloudSourceConstPtr src_upload_key_;
  PointCloudTargetConstPtr tgt_upload_key_;
  Matrix4 final_transformation_;
  ddlo_shim::Matrix6d final_hessian_;
  bool converged_ = false, lm_failed_ = false, have_correspondences_ = false, compute_output_ = true;
  int nr_iterations_ = 0;
};

}  // namespace nano_gicp
