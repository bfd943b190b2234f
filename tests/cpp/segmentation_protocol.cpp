// C++ host-side test of the DetectionModule shim (detection/detection.hpp): the calls OdomNode::applySegmentation
// makes (odom.cc:853-857) on an organised scan read from a file; the label image, the ground flags and the per-segment
// average residuals are written back for the Python harness to compare with the same stage driven through the C ABI.
//
//   segmentation_protocol in.bin out.bin
//   in.bin : int32 rows, cols, ground_rows; float32 ang_bottom, minimum_range, sensor_mount_angle, max_distance;
//            16 float32 pose (column-major); rows*cols*4 float32 scan (x y z 1, NaN = no return); rows*cols float32 residuals
//   out.bin: int32 label_count; rows*cols int32 labels; rows*cols int8 ground; label_count float64 avg residuals
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <detection/detection.hpp>

using PointType = ddlo_shim::PointXYZI;
using CloudT = ddlo_shim::Cloud<PointType>;

template <class T>
static void rd(FILE* f, T* p, size_t n) {
  if (std::fread(p, sizeof(T), n, f) != n) {
    std::fprintf(stderr, "short read\n");
    std::exit(2);
  }
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]);
    return 2;
  }
  try {
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int dims[3];
    float prm[4];
    ddlo_shim::Matrix4f T_;
    rd(f, dims, 3), rd(f, prm, 4), rd(f, T_.data(), 16);
    const int H = dims[0], W = dims[1];
    std::vector<float> scan((size_t)H * W * 4), res((size_t)H * W);
    rd(f, scan.data(), scan.size()), rd(f, res.data(), res.size());
    std::fclose(f);
    auto segmentation_scan_t_ = std::make_shared<CloudT>();
    auto residuals_cloud_ = std::make_shared<CloudT>();
    segmentation_scan_t_->points.resize((size_t)H * W);
    residuals_cloud_->points.resize((size_t)H * W);
    for (size_t i = 0; i < (size_t)H * W; ++i) {
      PointType& p = segmentation_scan_t_->points[i];
      p.x = scan[4 * i], p.y = scan[4 * i + 1], p.z = scan[4 * i + 2];
      residuals_cloud_->points[i].intensity = res[i];  // x = y = z = 0: finite, as the zero-filled cloud of odom.cc:807
    }
    // loadParams with this test's geometry; everything else keeps the reference's defaults
    ddlo_shim::DetectionModule defaults;
    (void)defaults;
    ddlo_segmentation_params p{};
    p.rows = H, p.cols = W, p.ground_rows = dims[2];
    p.valid_point_num = 15, p.min_line_num = 5, p.valid_line_num = 5;
    p.window_row_min = 0, p.window_row_max = H - 1, p.window_col_min = 0, p.window_col_max = W - 1;
    p.ang_bottom = prm[0], p.ground_angle_threshold = 10.0f, p.minimum_range = prm[1], p.sensor_mount_angle = prm[2];
    p.theta = static_cast<float>(60.0 / 180.0 * M_PI);
    p.min_delta_z = 0.1f, p.max_delta_z = 3.0f, p.max_distance = prm[3], p.max_elevation = 2.0f;
    ddlo_shim::DetectionModule detection_module_(&p);
    // OdomNode::applySegmentation (odom.cc:853-857)
    detection_module_.projectScan(segmentation_scan_t_, segmentation_scan_t_, T_, T_);
    detection_module_.projectResiduals(residuals_cloud_);
    detection_module_.applySegmentation();
    std::vector<int> ground_indices;
    detection_module_.getGroundIndices(ground_indices);
    size_t labelled = 0;
    for (const auto& v : detection_module_.label_indices_i_) labelled += v.size();
    std::printf("segments %zu ground %zu labelled %zu\n", detection_module_.getSegmentsCount(), ground_indices.size(), labelled);
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 2;
    std::fwrite(&detection_module_.label_count_, 4, 1, o);
    std::fwrite(detection_module_.label_mat_.data(), 4, (size_t)H * W, o);
    std::fwrite(detection_module_.ground_mat_.data(), 1, (size_t)H * W, o);
    std::fwrite(detection_module_.avg_residuals_.data(), 8, (size_t)detection_module_.label_count_, o);
    std::fclose(o);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
