// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// CPU restatement of the range-image segmentation stage of the reference's DetectionModule
// (SURVEY.md §8f row 4), sequential and in the reference's own visiting order:
//   projectScan        src/detection/detection.cpp:254-329   range image of the organised, world-frame scan
//   projectResiduals   src/detection/detection.cpp:203-252   residual image (organised copy)
//   groundRemoval      src/detection/detection.cpp:448-512   column-wise slope test on the lowest rows
//   cloudSegmentation  src/detection/detection.cpp:514-546   raster-order seeds
//   labelComponents    src/detection/detection.cpp:548-724   queue flood fill + segment feasibility
//
// PIN: the reference's detection.cpp as a whole needs ROS, OpenCV, PCL and the tracking module and cannot be compiled
// here, and the reference ships no fixtures for it.  Its class declaration (include/detection/detection.h, unmodified)
// and the definitions of exactly the member functions restated below are compiled from /root/reference over stand-in
// headers into oracle/_ref/libdetection_ref.so (ref_detection_shim.cpp, extract_detection.py, refdet.py);
// tests/test_reference_detection_cpu.py requires this file to reproduce that code bit for bit (labels, ground flags,
// ranges, average residuals), and the outputs of that code are committed as tests/golden/segmentation_reference.npz.
// Not covered by the pin: windows other than the reference's hard-coded 156..356 and the scan_in_sensor_frame option.
//
// Arithmetic notes (what the reference's build, -O2 without -march, evaluates):
//   * all members are float (include/detection/detection.h:60-86); products and sums are rounded one by one
//     (no FMA, this file is compiled with -ffp-contract=off);
//   * atan2 / sqrt / abs on float arguments resolve to the float overloads: detection.cpp sees `using namespace std;`
//     (detection.h -> tracking/tracking.h -> tracking/hungarian.h:42);
//   * groundRemoval keeps its loop temporaries outside the `omp parallel for` (a data race in the reference);
//     the sequential meaning is restated;
//   * the `valid_range` window is hard-coded to rows/cols 156..356 in the reference (:520-522, :565-567); here it
//     is a parameter with those defaults.  The window test sees the neighbour column BEFORE the wrap-around, as in
//     the reference (:590-599), so column -1 never passes and column W passes only if the window reaches W.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

extern "C" {

struct oracle_seg_params {  // same layout as ddlo_segmentation_params (include/ddlo_gicp.h)
  int rows, cols;
  int ground_rows;
  int valid_point_num, min_line_num, valid_line_num;
  int window_row_min, window_row_max, window_col_min, window_col_max;
  int scan_in_sensor_frame;
  int unordered_residual_sums;  // ignored here: the restatement always sums in push order
  float ang_bottom;
  float ground_angle_threshold, minimum_range, sensor_mount_angle, theta;
  float min_delta_z, max_delta_z, max_distance, max_elevation;
};

// scan_t: rows*cols points of `stride_floats` floats (x, y, z first), world frame, non-finite = no return.
// T16: column-major 4x4 float pose (Eigen::Matrix4f).  residuals: rows*cols floats (the intensity channel of the
// residual cloud) or null when projectResiduals was not called.
// Outputs: label_mat (int32), range_mat (float), ground_mat (int8), avg_residuals (double, rows*cols entries,
// index = label).  borderline[0] counts slope tests within a few ulp of their threshold (a different libm may
// decide those differently).  Returns label_count_ (first unused label; labels 1 .. label_count_-1 are segments).
int oracle_segment_scan(const oracle_seg_params* p, const float* scan_t, int stride_floats, const float* T16, const float* residuals,
                        int* label_mat, float* range_mat, signed char* ground_mat, double* avg_residuals, int* borderline) {
  const int H = p->rows, W = p->cols;
  const size_t HW = (size_t)H * W;
  const float nan = std::numeric_limits<float>::quiet_NaN();
  int n_border = 0;
  // loadParams (:78-79, :109-112)
  const float ang_res_x = 360.0 / float(W);
  const float ang_res_y = 2 * p->ang_bottom / float(H - 1);
  const float sin_ax = sin(ang_res_x / 180.0 * M_PI), cos_ax = cos(ang_res_x / 180.0 * M_PI);
  const float sin_ay = sin(ang_res_y / 180.0 * M_PI), cos_ay = cos(ang_res_y / 180.0 * M_PI);

  // resetParameters (:172-189)
  std::fill(label_mat, label_mat + HW, 0);
  std::fill(ground_mat, ground_mat + HW, (signed char)0);
  std::fill(range_mat, range_mat + HW, 0.0f);
  std::vector<float> full(HW * 3, nan);
  // OdomNode::transformScans (odom.cc:957-963) when the caller hands over the sensor-frame scan: pcl::transformPointCloud,
  // float (r0 x + r1 y) + (r2 z + t) as Eigen evaluates the 4x4 product; non-finite points are left alone
  std::vector<float> moved;
  if (p->scan_in_sensor_frame) {
    moved.resize(HW * 3);
    for (size_t i = 0; i < HW; ++i) {
      const float* q = scan_t + i * stride_floats;
      const float x = q[0], y = q[1], z = q[2];
      float* o = &moved[i * 3];
      o[0] = x, o[1] = y, o[2] = z;
      if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z))
        for (int r = 0; r < 3; ++r) o[r] = (T16[r] * x + T16[4 + r] * y) + (T16[8 + r] * z + T16[12 + r]);
    }
    scan_t = moved.data();
    stride_floats = 3;
  }
  auto pt = [&](size_t i) { return scan_t + i * stride_floats; };

  // projectScan (:296-327)
  const float x0 = -T16[12], y0 = -T16[13], z0 = -T16[14];
  for (int row = 0; row < H; ++row)
    for (int col = 0; col < W; ++col) {
      const size_t idx = (size_t)row * W + col;
      const float* q = pt(idx);
      if (!std::isfinite(q[0]) || !std::isfinite(q[1]) || !std::isfinite(q[2])) continue;
      const float x = q[0] + x0, y = q[1] + y0, z = q[2] + z0;
      const float range = sqrtf(x * x + y * y + z * z);
      if (range < p->minimum_range) continue;
      range_mat[idx] = range;
      full[idx * 3 + 0] = q[0], full[idx * 3 + 1] = q[1], full[idx * 3 + 2] = q[2];
    }

  // projectResiduals (:240-249): zero where the residual cloud has no finite point -- the caller passes the
  // intensity plane, already zero there
  std::vector<float> res(HW, 0.0f);
  if (residuals) std::copy(residuals, residuals + HW, res.begin());

  // groundRemoval (:460-510)
  for (int col = 0; col < W; ++col)
    for (int row_inverse = 0; row_inverse < p->ground_rows; ++row_inverse) {
      const int row = H - 1 - row_inverse;
      if (row - 1 < 0) break;  // the reference would index out of bounds here (ground_rows == H); not restated
      const size_t lower = (size_t)col + (size_t)row * W, upper = (size_t)col + (size_t)(row - 1) * W;
      if (full[lower * 3] == 0 || full[upper * 3] == 0) {
        ground_mat[lower] = -1;
        continue;
      }
      const float dx = full[upper * 3] - full[lower * 3];
      const float dy = full[upper * 3 + 1] - full[lower * 3 + 1];
      const float dz = full[upper * 3 + 2] - full[lower * 3 + 2];
      const float angle = std::atan2(dz, std::sqrt(dx * dx + dy * dy)) * 180 / M_PI;
      const float dev = std::fabs(angle - p->sensor_mount_angle);
      if (std::fabs(dev - p->ground_angle_threshold) <= 1e-4f) ++n_border;
      if (dev <= p->ground_angle_threshold) {
        ground_mat[lower] = 1;
        ground_mat[upper] = 1;
      }
    }
  for (size_t i = 0; i < HW; ++i)
    if (ground_mat[i] == 1 || range_mat[i] == 0) label_mat[i] = -1;

  // cloudSegmentation + labelComponents (:514-724)
  auto in_window = [&](size_t i, size_t j) {  // size_t on purpose: a column of -1 becomes huge, as in the reference
    return i >= (size_t)p->window_row_min && i <= (size_t)p->window_row_max && j >= (size_t)p->window_col_min &&
           j <= (size_t)p->window_col_max;
  };
  static const int kDy[4] = {-1, 0, 0, 1};
  static const int kDx[4] = {0, 1, -1, 0};
  std::vector<int> queue_y(HW), queue_x(HW);
  std::vector<char> line_flag(H);
  const float current_height = T16[14];
  int label_count = 1;
  for (int si = 0; si < H; ++si)
    for (int sj = 0; sj < W; ++sj) {
      if (label_mat[(size_t)si * W + sj] != 0 || !in_window(si, sj)) continue;
      std::fill(line_flag.begin(), line_flag.end(), 0);
      queue_y[0] = si, queue_x[0] = sj;
      int head = 0, tail = 1;
      float min_z = 1e6, max_z = -1e6, min_dist = 1e6, max_dist = -1e6, total_residuum = 0;
      int res_count = 0;
      while (head < tail) {
        const int fy = queue_y[head], fx = queue_x[head];
        ++head;
        label_mat[(size_t)fy * W + fx] = label_count;
        for (int n = 0; n < 4; ++n) {
          const int ty = fy + kDy[n];
          int tx = fx + kDx[n];
          if (ty < 0 || ty >= H) continue;
          if (!in_window(ty, tx)) continue;
          if (tx < 0) tx = W - 1;
          if (tx >= W) tx = 0;
          const size_t ti = (size_t)ty * W + tx;
          if (label_mat[ti] != 0) continue;
          const float rf = range_mat[(size_t)fy * W + fx], rt = range_mat[ti];
          const float d1 = std::max(rf, rt), d2 = std::min(rf, rt);
          const float sin_a = kDy[n] == 0 ? sin_ax : sin_ay, cos_a = kDy[n] == 0 ? cos_ax : cos_ay;
          const float angle = std::atan2(d2 * sin_a, (d1 - d2 * cos_a));
          if (std::fabs(angle - p->theta) <= 1e-6f) ++n_border;
          if (angle > p->theta) {
            const double z = pt(ti)[2];
            if (z < min_z && z != 0)
              min_z = z;
            else if (z > max_z)
              max_z = z;
            min_dist = std::min(min_dist, std::min(d1, d2));
            max_dist = std::max(max_dist, std::max(d1, d2));
            queue_y[tail] = ty, queue_x[tail] = tx;
            ++tail;
            label_mat[ti] = label_count;
            line_flag[ty] = 1;
            if (res[ti] > 0) {
              total_residuum += res[ti];
              ++res_count;
            }
          }
        }
      }
      const int pushed = tail;  // all_pushed_ind_size: the seed and every queued pixel
      int line_count = 0;
      for (int i = 0; i < H; ++i) line_count += line_flag[i] ? 1 : 0;
      bool feasible = false;
      if (pushed >= 50 && line_count >= p->min_line_num)
        feasible = true;
      else if (pushed >= p->valid_point_num && line_count >= p->valid_line_num)
        feasible = true;
      if (feasible) feasible = max_dist <= p->max_distance;
      if (feasible) {
        const float delta_z = max_z - min_z;
        feasible = p->min_delta_z <= delta_z && delta_z <= p->max_delta_z;
      }
      if (feasible) feasible = min_z - current_height <= p->max_elevation;
      if (feasible) {
        double avg = 0;
        if (residuals) avg = res_count > 0 ? total_residuum / res_count : 0;
        avg_residuals[label_count] = avg;
        ++label_count;
      } else {
        for (int i = 0; i < pushed; ++i) label_mat[(size_t)queue_y[i] * W + queue_x[i]] = 999999;
      }
    }
  if (borderline) *borderline = n_border;
  return label_count;
}

}  // extern "C"
