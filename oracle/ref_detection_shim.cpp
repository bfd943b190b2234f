// ORACLE / TEST INFRASTRUCTURE ONLY -- the reference's OWN DetectionModule code on the segmentation path, compiled from
// where it lies under /root/reference, behind a small C interface for ctypes (oracle/refdet.py).
//
//   * the class declaration is the reference's include/detection/detection.h, unmodified (its single include,
//     <tracking/tracking.h>, resolves to the stand-in of oracle/stub_include/);
//   * the definitions of loadParams, allocateMemory, resetParameters, projectResiduals, projectScan, groundRemoval,
//     cloudSegmentation and labelComponents are the reference's, extracted from src/detection/detection.cpp by
//     oracle/extract_detection.py into _ref/detection_extract.inc at build time;
//   * the constructor below is ours (the reference's also sets up ROS publishers and the evaluation).
// The window of the `valid_range` lambda (rows/cols 156..356) and the int-typed ROS parameters (ang_bottom,
// groundAngleThreshold, minimumRange, sensorMountAngle, maxDistance take the type of their integer defaults) are the
// reference's: cases for this library use images larger than 356 pixels and integer values for those parameters.
// groundRemoval's `omp parallel for` shares its loop temporaries between threads (a data race in the reference): this
// library runs it with one thread.
#include <omp.h>

#define private public
#include <detection/detection.h>
#undef private

typedef pcl::PointXYZI PointType;  // odometry/ddlo.h:90, for the loop taken from odom.cc

#include "_ref/detection_extract.inc"

DetectionModule::DetectionModule() : initialized_(false), icp_residuals_set_(false), it_(nh_) {
  loadParams();
  allocateMemory();
  resetParameters();
}

extern "C" {

// names/values: ROS parameter overrides ("odomNode/detection/rows", ...).  scan / scan_t: rows*cols points (x y z w floats),
// sensor and world frame (only the second enters the results).  T16, T_s2s16: column-major.  residual_intensity: rows*cols
// floats or null (projectResiduals not called).  Returns label_count_, or -1 if the image size does not match.
int refdet_segment(const char* const* names, const double* values, int n_params, const float* scan, const float* scan_t, int n_points,
                   const float* T16, const float* residual_intensity, int* label_mat, float* range_mat, signed char* ground_mat,
                   double* avg_residuals, int avg_capacity) {
  omp_set_num_threads(1);
  ddlo_refdet::overrides().clear();
  for (int i = 0; i < n_params; ++i) ddlo_refdet::overrides()[names[i]] = values[i];
  DetectionModule det;
  const int H = det.H_, W = det.W_;
  if (n_points != H * W) return -1;
  using Cloud = pcl::PointCloud<pcl::PointXYZI>;
  Cloud::Ptr cloud_in(new Cloud), cloud_in_t(new Cloud), residuals(new Cloud);
  for (Cloud::Ptr* c : {&cloud_in, &cloud_in_t, &residuals}) {
    (*c)->points.resize((size_t)n_points);
    (*c)->width = W, (*c)->height = H;
  }
  for (int i = 0; i < n_points; ++i) {
    cloud_in->points[i].x = scan[4 * i], cloud_in->points[i].y = scan[4 * i + 1], cloud_in->points[i].z = scan[4 * i + 2];
    cloud_in_t->points[i].x = scan_t[4 * i], cloud_in_t->points[i].y = scan_t[4 * i + 1], cloud_in_t->points[i].z = scan_t[4 * i + 2];
    if (residual_intensity) residuals->points[i].intensity = residual_intensity[i];
  }
  Eigen::Matrix4f T;
  std::memcpy(T.m, T16, sizeof(T.m));
  // OdomNode::applySegmentation (odom.cc:853-857) up to cloudSegmentation
  det.projectScan(cloud_in, cloud_in_t, T, T);
  if (residual_intensity)
    det.projectResiduals(residuals);
  else  // labelComponents reads residuals_mat_ even when icp_residuals_set_ is false (it then holds the previous frame's image,
        // or nothing at all on the first frame); its content cannot reach an output in that case, any image of the right size will do
    det.residuals_mat_ = cv::Mat(H, W, CV_32F, cv::Scalar::all(0));
  det.groundRemoval();
  det.cloudSegmentation();
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < W; ++c) {
      label_mat[r * W + c] = det.label_mat_.at<int>(r, c);
      range_mat[r * W + c] = det.range_mat_.at<float>(r, c);
      ground_mat[r * W + c] = det.ground_mat_.at<int8_t>(r, c);
    }
  for (int i = 0; i < det.label_count_ && i < avg_capacity; ++i) avg_residuals[i] = det.avg_residuals_[i];
  return det.label_count_;
}

// The residual cloud of OdomNode::scanMatching (odom.cc:804-827), the reference's own loop: 512 x 512 cells, +-60 degrees.
// scan: n points (x y z w floats); residuals: n doubles (getResiduals).  out_xyzi: 512*512*4 floats (x, y, z, intensity).
void refdet_residual_cloud(const float* scan, int n, const double* residuals_in, float* out_xyzi) {
  using Cloud = pcl::PointCloud<pcl::PointXYZI>;
  Cloud::Ptr registration_scan_(new Cloud), residuals_cloud_(new Cloud);
  registration_scan_->points.resize((size_t)n);
  for (int i = 0; i < n; ++i) {
    registration_scan_->points[i].x = scan[4 * i], registration_scan_->points[i].y = scan[4 * i + 1], registration_scan_->points[i].z = scan[4 * i + 2];
  }
  std::vector<double> residuals(residuals_in, residuals_in + n);
  ref_residual_cloud_body(registration_scan_, residuals, residuals_cloud_);
  for (size_t c = 0; c < residuals_cloud_->points.size(); ++c) {
    const pcl::PointXYZI& p = residuals_cloud_->points[c];
    out_xyzi[4 * c] = p.x, out_xyzi[4 * c + 1] = p.y, out_xyzi[4 * c + 2] = p.z, out_xyzi[4 * c + 3] = p.intensity;
  }
}

}  // extern "C"
