// Stand-in for pcl/registration/registration.h (TEST INFRASTRUCTURE): the part of pcl::Registration<Source, Target,
// float> (PCL 1.10) that nano_gicp::LsqRegistration / NanoGICP derive from and OdomNode calls: input/target
// pointers, the iteration / epsilon / distance knobs with PCL's defaults, align() -> computeTransformation(),
// getFinalTransformation(), hasConverged(), and pcl::transformPointCloud (float arithmetic).
#ifndef DDLO_ORACLE_PCL_REGISTRATION_STUB
#define DDLO_ORACLE_PCL_REGISTRATION_STUB
#include <cmath>
#include <limits>
#include <string>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace pcl {
template <class PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  if (&in != &out) out = in;
  for (std::size_t i = 0; i < out.points.size(); ++i) {
    const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    out.points[i].x = T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3);
    out.points[i].y = T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3);
    out.points[i].z = T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3);
  }
}
template <class PointSource, class PointTarget, class Scalar = float>
class Registration {
 public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using PointCloudSource = PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  Registration()
      : nr_iterations_(0), max_iterations_(10), final_transformation_(Matrix4::Identity()), transformation_epsilon_(0.0),
        corr_dist_threshold_(std::sqrt(std::numeric_limits<double>::max())), converged_(false) {}
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  PointCloudSourceConstPtr getInputSource() const { return input_; }
  PointCloudTargetConstPtr getInputTarget() const { return target_; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  void setEuclideanFitnessEpsilon(double) {}
  void setRANSACIterations(int) {}
  void setRANSACOutlierRejectionThreshold(double) {}
  Matrix4 getFinalTransformation() const { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (!input_ || !target_) return;  // initCompute() fails: PCL returns without touching anything
    output.points.resize(input_->points.size());
    converged_ = false;
    final_transformation_ = Matrix4::Identity();
    computeTransformation(output, guess);
  }

 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_;
  int max_iterations_;
  Matrix4 final_transformation_;
  double transformation_epsilon_;
  double corr_dist_threshold_;
  bool converged_;
};
}  // namespace pcl
#endif
