// Stand-in for <tracking/tracking.h> (TEST INFRASTRUCTURE).  The reference's include/detection/detection.h includes only
// this header and gets ROS, OpenCV, PCL, Eigen and the tracking module through it; none of them is installed here.  This
// file provides just the types and calls that the class declaration of DetectionModule and its eight member functions on
// the segmentation path (oracle/extract_detection.py) touch, so that the reference's own text compiles unmodified:
// cv::Mat / Scalar / Vec3b, pcl::PointXYZI / PointCloud / isFinite / copyPointCloud, Eigen::Matrix4f::coeff,
// ros::NodeHandle::param (with an override table the shim fills), the ROS_* log macros, and empty shells for the rest.
// The real tracking.h pulls in tracking/hungarian.h, whose `using namespace std;` (hungarian.h:42) makes the unqualified
// abs / atan2 / sqrt calls of the reference resolve to the float overloads; the same directive is repeated here (and
// <math.h> / <stdlib.h> are included, which has the same effect with libstdc++).
#ifndef DDLO_ORACLE_TRACKING_STUB
#define DDLO_ORACLE_TRACKING_STUB
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

using namespace std;  // tracking/hungarian.h:42

#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)

namespace Eigen {
struct Matrix4f {  // column-major, like Eigen
  float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  float coeff(int r, int c) const { return m[c * 4 + r]; }
};
}  // namespace Eigen

namespace pcl {
struct PCLHeader {};
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, data3 = 1.0f;
  float intensity = 0, pad[3] = {0, 0, 0};
};
template <class PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  void resize(std::size_t n) { points.resize(n); }
  void clear() {
    points.clear();
    width = height = 0;
  }
  bool isOrganized() const { return height > 1; }
};
inline bool isFinite(const PointXYZI& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }
template <class PointT>
void copyPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out) {
  out = in;
}
}  // namespace pcl

namespace cv {
struct Scalar {
  double v;
  static Scalar all(double x) { return Scalar{x}; }
};
struct Vec3b {
  unsigned char v[3] = {0, 0, 0};
  Vec3b() = default;
  Vec3b(int a, int b, int c = 0) : v{(unsigned char)a, (unsigned char)b, (unsigned char)c} {}
};
enum { DDLO_CV_8S = 1, DDLO_CV_32S = 4, DDLO_CV_32F = 5 };
#define CV_8S cv::DDLO_CV_8S
#define CV_32S cv::DDLO_CV_32S
#define CV_32F cv::DDLO_CV_32F
class Mat {  // dense, row-major, one channel; every element set to the scalar at construction
 public:
  int rows = 0, cols = 0, type = 0;
  std::vector<unsigned char> bytes;
  Mat() = default;
  Mat(int r, int c, int t, const Scalar& s) : rows(r), cols(c), type(t), bytes((std::size_t)r * c * elem(t)) {
    for (std::size_t i = 0; i < (std::size_t)r * c; ++i) {
      if (t == DDLO_CV_8S) reinterpret_cast<signed char*>(bytes.data())[i] = (signed char)s.v;
      if (t == DDLO_CV_32S) reinterpret_cast<int*>(bytes.data())[i] = (int)s.v;
      if (t == DDLO_CV_32F) reinterpret_cast<float*>(bytes.data())[i] = (float)s.v;
    }
  }
  static int elem(int t) { return t == DDLO_CV_8S ? 1 : 4; }
  template <class T>
  T& at(int r, int c) {
    return reinterpret_cast<T*>(bytes.data())[(std::size_t)r * cols + c];
  }
};
}  // namespace cv

namespace ddlo_refdet {  // parameter overrides the shim sets before constructing the module ("odomNode/detection/rows" -> value)
inline std::map<std::string, double>& overrides() {
  static std::map<std::string, double> m;
  return m;
}
}  // namespace ddlo_refdet

namespace ros {
struct Time {
  Time() = default;
  explicit Time(double) {}
};
struct NodeHandle {
  // ros::NodeHandle::param(name, default): the value's C++ type is the type of the default, as in roscpp
  template <class T>
  T param(const std::string& name, const T& def) const {
    if constexpr (std::is_arithmetic_v<T>) {
      auto it = ddlo_refdet::overrides().find(name);
      if (it != ddlo_refdet::overrides().end()) return static_cast<T>(it->second);
    }
    return def;
  }
};
}  // namespace ros

namespace image_transport {
struct Publisher {};
struct ImageTransport {
  explicit ImageTransport(const ros::NodeHandle&) {}
};
}  // namespace image_transport

namespace std_msgs {
struct Header {};
}  // namespace std_msgs
namespace pcl_conversions {
inline void fromPCL(const pcl::PCLHeader&, std_msgs::Header&) {}
}  // namespace pcl_conversions

struct AccumulatorData {  // util/accumulator.h: wall-clock statistics, no influence on results
  void tick() {}
  void tock() {}
};
enum class ObjectStatus { UNDEFINED, STATIC, DYNAMIC };
struct Object {};
struct TrackingModule {};
#endif
