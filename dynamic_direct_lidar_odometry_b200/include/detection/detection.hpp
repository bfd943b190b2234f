// DetectionModule's segmentation stage on the B200 — header-only C++17 shim over the C ABI (include/ddlo_gicp.h).
//
// Mirrors the part of the reference's DetectionModule that OdomNode::applySegmentation drives every frame
//   /root/reference/dynamic_direct_lidar_odometry/include/detection/detection.h:20-150
//   /root/reference/dynamic_direct_lidar_odometry/src/detection/detection.cpp:72-126   loadParams
//                                                                           :203-252  projectResiduals
//                                                                           :254-329  projectScan
//                                                                           :191-196  applySegmentation -> groundRemoval, cloudSegmentation
// with the same method and member names (H_, W_, label_mat_, range_mat_, ground_mat_, avg_residuals_, label_count_,
// label_indices_i_, icp_residuals_set_), so that the rest of the reference's module (computeAllObjects, trackDetections,
// visualize, the getters) keeps reading what it read before.  The images are plain row-major vectors here instead of
// cv::Mat (OpenCV is not in this repository's build image); with OpenCV the same buffers wrap into cv::Mat headers
// without a copy.  All arithmetic runs in libddlo_gicp_b200.so on the GPU (ddlo_segment_scan / ddlo_gicp_segment_scan).
#pragma once

#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../nano_gicp/nano_gicp.hpp"

namespace ddlo_shim {

class DetectionModule {
public:
  using PointType = PointXYZI;
  using CloudPtr = typename Cloud<PointType>::Ptr;

  // loadParams (:72-105): the ROS parameters of odomNode/detection/*, same defaults; the valid_range window the
  // reference hard-codes (:520-522) is a parameter here
  explicit DetectionModule(const ddlo_segmentation_params* params = nullptr, ddlo_runtime* rt = nullptr) : rt_(rt ? rt : runtime()) {
    params_.rows = 128, params_.cols = 1024, params_.ground_rows = 30;
    params_.valid_point_num = 15, params_.min_line_num = 5, params_.valid_line_num = 5;
    params_.window_row_min = 156, params_.window_row_max = 356, params_.window_col_min = 156, params_.window_col_max = 356;
    params_.scan_in_sensor_frame = 0, params_.unordered_residual_sums = 0;
    params_.ang_bottom = 45.0f, params_.ground_angle_threshold = 10.0f, params_.minimum_range = 10.0f, params_.sensor_mount_angle = 10.0f;
    params_.theta = static_cast<float>(60.0 / 180.0 * M_PI);
    params_.min_delta_z = 0.1f, params_.max_delta_z = 3.0f, params_.max_distance = 20.0f, params_.max_elevation = 2.0f;
    if (params) params_ = *params;
    H_ = params_.rows, W_ = params_.cols;
    avg_residuals_.assign((size_t)H_ * W_, 0.0);  // allocateMemory (:162)
    resetParameters();
  }

  // :254-329 — only the transformed scan and the pose enter the range image
  // (a null cloud_in_t: the sensor-frame scan cloud_in is moved by T on the device, OdomNode::transformScans, odom.cc:957-963)
  void projectScan(const CloudPtr& cloud_in, const CloudPtr& cloud_in_t, const Matrix4f& T, const Matrix4f& T_s2s = identity4f()) {
    const CloudPtr& scan = cloud_in_t ? cloud_in_t : cloud_in;
    if (!scan || scan->size() != (size_t)H_ * W_) throw std::invalid_argument("DetectionModule: the segmentation scan must be organised (H_ x W_ points)");
    resetParameters();
    params_.scan_in_sensor_frame = cloud_in_t ? 0 : 1;
    cloud_in_t_ = scan;
    T_ = T, T_s2s_ = T_s2s;
  }

  // :203-252 — residual image = intensity of the organised residual cloud where its point is finite
  void projectResiduals(const CloudPtr& cloud_in) {
    if (!cloud_in || cloud_in->size() != (size_t)H_ * W_) throw std::invalid_argument("DetectionModule: the residual cloud must have H_ x W_ points");
    residuals_mat_.assign((size_t)H_ * W_, 0.0f);
    for (size_t i = 0; i < residuals_mat_.size(); ++i) {
      const PointType& pt = cloud_in->points[i];
      if (std::isfinite(pt.x) && std::isfinite(pt.y) && std::isfinite(pt.z)) residuals_mat_[i] = pt.intensity;
    }
    engine_ = nullptr;
    icp_residuals_set_ = true;
  }
  // the same without the host round trip: the residual cloud of `engine`'s last align (odom.cc:804-827) is built and read on the device
  template <class Engine>
  void projectResidualsFrom(const Engine& engine, double angle_min = -M_PI / 3, double angle_max = M_PI / 3) {
    engine_ = engine.handle();
    angle_min_ = angle_min, angle_max_ = angle_max;
    icp_residuals_set_ = true;
  }

  // :191-199 — groundRemoval() + cloudSegmentation(); the label indices of :528-543 are rebuilt from the label image
  void applySegmentation() {
    if (!cloud_in_t_) throw std::logic_error("DetectionModule: projectScan has not been called");
    const float* scan = &cloud_in_t_->points[0].x;
    int rc;
    if (icp_residuals_set_ && engine_)
      rc = ddlo_gicp_segment_scan(engine_, &params_, scan, (int)sizeof(PointType), data(T_), angle_min_, angle_max_, label_mat_.data(),
                                  range_mat_.data(), ground_mat_.data(), avg_residuals_.data(), (int)avg_residuals_.size(), &label_count_, nullptr);
    else
      rc = ddlo_segment_scan(rt_, &params_, scan, (int)sizeof(PointType), data(T_), icp_residuals_set_ ? residuals_mat_.data() : nullptr,
                             label_mat_.data(), range_mat_.data(), ground_mat_.data(), avg_residuals_.data(), (int)avg_residuals_.size(),
                             &label_count_, nullptr);
    if (rc != DDLO_OK) throw std::runtime_error(std::string("ddlo_segment_scan: ") + ddlo_last_error());
    label_indices_i_.assign(label_count_, {});
    for (int i = 0; i < H_ * W_; ++i) {
      const int label = label_mat_[i];
      if (label > 0 && label != 999999) label_indices_i_[label].push_back(i);
    }
    icp_residuals_set_ = false;
    engine_ = nullptr;
  }

  size_t getSegmentsCount() const { return (size_t)label_count_ - 1; }
  void getGroundIndices(std::vector<int>& indices) const {
    indices.clear();
    for (int i = 0; i < H_ * W_; ++i)
      if (ground_mat_[i] == 1) indices.push_back(i);
  }

  int H_ = 0, W_ = 0;
  int label_count_ = 1;
  bool icp_residuals_set_ = false;
  std::vector<int> label_mat_;            // H_ x W_, row-major (cv::Mat CV_32S in the reference)
  std::vector<float> range_mat_;          // CV_32F
  std::vector<signed char> ground_mat_;   // CV_8S
  std::vector<float> residuals_mat_;      // CV_32F
  std::vector<double> avg_residuals_;     // indexed by label
  std::vector<std::vector<int>> label_indices_i_;

private:
  void resetParameters() {  // :172-189
    label_count_ = 1;
    label_indices_i_.clear();
    label_mat_.assign((size_t)H_ * W_, 0);
    ground_mat_.assign((size_t)H_ * W_, 0);
    range_mat_.assign((size_t)H_ * W_, 0.0f);
  }

  ddlo_runtime* rt_ = nullptr;
  ddlo_segmentation_params params_{};
  CloudPtr cloud_in_t_;
  Matrix4f T_ = identity4f(), T_s2s_ = identity4f();
  ddlo_gicp* engine_ = nullptr;
  double angle_min_ = 0, angle_max_ = 0;
};

}  // namespace ddlo_shim
