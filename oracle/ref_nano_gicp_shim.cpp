// TEST INFRASTRUCTURE.  Builds the reference's OWN nano_gicp engine,
//   /root/reference/dynamic_direct_lidar_odometry/include/nano_gicp/{nano_gicp,lsq_registration,nanoflann}.hpp
//   + impl/{nano_gicp_impl,lsq_registration_impl,nanoflann_impl}.hpp + gicp/{so3,gicp_settings}.hpp,
// unmodified and from where the files lie, into oracle/_ref/libnano_gicp_ref.so, against the stand-in headers for
// Eigen / PCL / Boost in oracle/stub_include (none of the three exists in this image; see Eigen/Core there for what
// the stand-in restates).  The explicit instantiation mirrors src/nano_gicp/nano_gicp.cc of the reference.
// A flat C interface (used by oracle/pyoracle.py through ctypes) drives the engine the way OdomNode does.
// Nothing of the product links or loads this file; tests use it to check the oracle's restatement
// (oracle_gicp.cpp) against the reference's control flow and formulas as written.
#include <omp.h>  // (in the reference build it arrives through the PCL headers)

#include <cstring>
#include <memory>
#include <vector>

#include <nano_gicp/nano_gicp.hpp>
#include <nano_gicp/impl/lsq_registration_impl.hpp>
#include <nano_gicp/impl/nano_gicp_impl.hpp>

namespace {
using Point = pcl::PointXYZI;
using Cloud = pcl::PointCloud<Point>;
using Covs = std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>;

// opens the protected hooks of the reference class for the stepwise parity tests
class Engine : public nano_gicp::NanoGICP<Point, Point> {
 public:
  using Base = nano_gicp::NanoGICP<Point, Point>;
  double call_linearize(const Eigen::Isometry3d& T, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) { return this->linearize(T, H, b); }
  double call_compute_error(const Eigen::Isometry3d& T) { return this->compute_error(T); }
  int nr_iterations() const { return this->nr_iterations_; }
  const std::vector<int>& corr() const { return this->correspondences_; }
  const std::vector<float>& sqd() const { return this->sq_distances_; }
  const Covs& mahal() const { return this->mahalanobis_; }
  void set_optimizer(int t) { this->lsq_optimizer_type_ = t == 0 ? nano_gicp::LSQ_OPTIMIZER_TYPE::GaussNewton : nano_gicp::LSQ_OPTIMIZER_TYPE::LevenbergMarquardt; }
  void set_lm_max_iterations(int n) { this->lm_max_iterations_ = n; }
};

Cloud::Ptr make_cloud(const float* xyz, int n, int stride_floats) {
  Cloud::Ptr c(new Cloud);
  c->resize((size_t)n);
  for (int i = 0; i < n; ++i) {
    c->points[i].x = xyz[(size_t)i * stride_floats + 0];
    c->points[i].y = xyz[(size_t)i * stride_floats + 1];
    c->points[i].z = xyz[(size_t)i * stride_floats + 2];
  }
  return c;
}
Eigen::Isometry3d iso_from_colmajor(const double* m) {
  Eigen::Matrix4d M;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) M(i, j) = m[4 * j + i];
  return Eigen::Isometry3d(M);
}
void covs_out(const Covs& v, double* out) {
  for (size_t k = 0; k < v.size(); ++k)
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) out[16 * k + 4 * j + i] = v[k](i, j);
}
Covs covs_in(const double* m16, int n) {
  Covs v((size_t)n);
  for (int k = 0; k < n; ++k)
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) v[k](i, j) = m16[16 * (size_t)k + 4 * j + i];
  return v;
}
}  // namespace

template class nano_gicp::NanoGICP<pcl::PointXYZI, pcl::PointXYZI>;
template class nano_gicp::LsqRegistration<pcl::PointXYZI, pcl::PointXYZI>;

extern "C" {

void* refgicp_cloud_create(const float* xyz, int n, int stride_floats) { return new Cloud::Ptr(make_cloud(xyz, n, stride_floats)); }
void refgicp_cloud_free(void* c) { delete static_cast<Cloud::Ptr*>(c); }

void* refgicp_create() { return new Engine(); }
void refgicp_free(void* g) { delete static_cast<Engine*>(g); }
void refgicp_set_num_threads(void* g, int n) { static_cast<Engine*>(g)->setNumThreads(n); }
void refgicp_set_correspondence_randomness(void* g, int k) { static_cast<Engine*>(g)->setCorrespondenceRandomness(k); }
void refgicp_set_regularization_method(void* g, int m) { static_cast<Engine*>(g)->setRegularizationMethod(static_cast<nano_gicp::RegularizationMethod>(m)); }
void refgicp_set_max_correspondence_distance(void* g, double d) { static_cast<Engine*>(g)->setMaxCorrespondenceDistance(d); }
void refgicp_set_maximum_iterations(void* g, int n) { static_cast<Engine*>(g)->setMaximumIterations(n); }
void refgicp_set_transformation_epsilon(void* g, double e) { static_cast<Engine*>(g)->setTransformationEpsilon(e); }
void refgicp_set_rotation_epsilon(void* g, double e) { static_cast<Engine*>(g)->setRotationEpsilon(e); }
void refgicp_set_initial_lambda_factor(void* g, double f) { static_cast<Engine*>(g)->setInitialLambdaFactor(f); }
void refgicp_set_lm_max_iterations(void* g, int n) { static_cast<Engine*>(g)->set_lm_max_iterations(n); }
void refgicp_set_optimizer(void* g, int t) { static_cast<Engine*>(g)->set_optimizer(t); }

void refgicp_set_input_source(void* g, void* c) { static_cast<Engine*>(g)->setInputSource(*static_cast<Cloud::Ptr*>(c)); }
void refgicp_set_input_target(void* g, void* c) { static_cast<Engine*>(g)->setInputTarget(*static_cast<Cloud::Ptr*>(c)); }
void refgicp_register_input_source(void* g, void* c) { static_cast<Engine*>(g)->registerInputSource(*static_cast<Cloud::Ptr*>(c)); }
void refgicp_clear_source(void* g) { static_cast<Engine*>(g)->clearSource(); }
void refgicp_clear_target(void* g) { static_cast<Engine*>(g)->clearTarget(); }
void refgicp_swap_source_and_target(void* g) { static_cast<Engine*>(g)->swapSourceAndTarget(); }
// `s2m.source_kdtree_ = s2s.source_kdtree_; s2m.source_covs_.clear();` (odom.cc:530-531)
void refgicp_share_source_tree(void* dst, void* src) {
  static_cast<Engine*>(dst)->source_kdtree_ = static_cast<Engine*>(src)->source_kdtree_;
  static_cast<Engine*>(dst)->source_covs_.clear();
}
int refgicp_calculate_source_covariances(void* g) { return static_cast<Engine*>(g)->calculateSourceCovariances() ? 0 : -1; }
int refgicp_calculate_target_covariances(void* g) { return static_cast<Engine*>(g)->calculateTargetCovariances() ? 0 : -1; }
int refgicp_source_covs_size(void* g) { return (int)static_cast<Engine*>(g)->getSourceCovariances().size(); }
int refgicp_target_covs_size(void* g) { return (int)static_cast<Engine*>(g)->getTargetCovariances().size(); }
void refgicp_get_source_covariances(void* g, double* out) { covs_out(static_cast<Engine*>(g)->getSourceCovariances(), out); }
void refgicp_get_target_covariances(void* g, double* out) { covs_out(static_cast<Engine*>(g)->getTargetCovariances(), out); }
void refgicp_set_source_covariances(void* g, const double* m16, int n) { static_cast<Engine*>(g)->setSourceCovariances(covs_in(m16, n)); }
void refgicp_set_target_covariances(void* g, const double* m16, int n) { static_cast<Engine*>(g)->setTargetCovariances(covs_in(m16, n)); }

// align(output, guess): final transformation (column-major float 4x4), converged_, nr_iterations_, final hessian
int refgicp_align(void* g, const float* guess16, float* final16, int* converged, int* nr_iterations, double* hessian36) {
  Engine* e = static_cast<Engine*>(g);
  Eigen::Matrix4f G = Eigen::Matrix4f::Identity();
  if (guess16)
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 4; ++i) G(i, j) = guess16[4 * j + i];
  Cloud out;
  e->align(out, G);
  const Eigen::Matrix4f T = e->getFinalTransformation();
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) final16[4 * j + i] = T(i, j);
  *converged = e->hasConverged() ? 1 : 0;
  *nr_iterations = e->nr_iterations();
  const Eigen::Matrix<double, 6, 6>& H = e->getFinalHessian();
  for (int j = 0; j < 6; ++j)
    for (int i = 0; i < 6; ++i) hessian36[6 * j + i] = H(i, j);
  return 0;
}
int refgicp_linearize(void* g, const double* T16, double* H36, double* b6, double* err) {
  Engine* e = static_cast<Engine*>(g);
  Eigen::Matrix<double, 6, 6> H;
  Eigen::Matrix<double, 6, 1> b;
  *err = e->call_linearize(iso_from_colmajor(T16), &H, &b);
  for (int j = 0; j < 6; ++j)
    for (int i = 0; i < 6; ++i) H36[6 * j + i] = H(i, j);
  for (int i = 0; i < 6; ++i) b6[i] = b(i);
  return 0;
}
int refgicp_compute_error(void* g, const double* T16, double* err) {
  *err = static_cast<Engine*>(g)->call_compute_error(iso_from_colmajor(T16));
  return 0;
}
int refgicp_get_correspondences(void* g, int* corr, float* sqd) {
  Engine* e = static_cast<Engine*>(g);
  std::memcpy(corr, e->corr().data(), e->corr().size() * sizeof(int));
  std::memcpy(sqd, e->sqd().data(), e->sqd().size() * sizeof(float));
  return (int)e->corr().size();
}
int refgicp_get_mahalanobis(void* g, double* out16) {
  covs_out(static_cast<Engine*>(g)->mahal(), out16);
  return (int)static_cast<Engine*>(g)->mahal().size();
}
int refgicp_get_residuals(void* g, double* out) {
  std::vector<double> r;
  static_cast<Engine*>(g)->getResiduals(r, Eigen::Matrix4f::Identity());
  std::memcpy(out, r.data(), r.size() * sizeof(double));
  return (int)r.size();
}
int refgicp_get_residual_vectors(void* g, const float* T16, float* out3) {
  Eigen::Matrix4f T;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) T(i, j) = T16[4 * j + i];
  std::vector<Eigen::Vector3f> r;
  static_cast<Engine*>(g)->getResiduals(r, T);
  for (size_t k = 0; k < r.size(); ++k)
    for (int i = 0; i < 3; ++i) out3[3 * k + i] = r[k](i);
  return (int)r.size();
}

}  // extern "C"
