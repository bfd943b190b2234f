// Stand-in for pcl/point_types.h (TEST INFRASTRUCTURE): pcl::PointXYZI with PCL's layout (32 bytes: x y z 1 |
// intensity + padding) and the getVector4fMap() view the reference's nano_gicp sources use.
#ifndef DDLO_ORACLE_PCL_POINT_TYPES_STUB
#define DDLO_ORACLE_PCL_POINT_TYPES_STUB
#include <Eigen/Core>
namespace pcl {
// what Eigen::Map<Eigen::Vector4f, Eigen::Aligned> is used for: read as a vector, assign from a vector
struct Vector4fMap {
  float* p;
  operator Eigen::Vector4f() const { return Eigen::Vector4f(p[0], p[1], p[2], p[3]); }
  Vector4fMap& operator=(const Eigen::Vector4f& v) {
    for (int i = 0; i < 4; ++i) p[i] = v(i);
    return *this;
  }
  template <class U>
  Eigen::Matrix<U, 4, 1> cast() const {
    return Eigen::Matrix<U, 4, 1>(static_cast<U>(p[0]), static_cast<U>(p[1]), static_cast<U>(p[2]), static_cast<U>(p[3]));
  }
};
struct Vector4fMapConst {
  const float* p;
  operator Eigen::Vector4f() const { return Eigen::Vector4f(p[0], p[1], p[2], p[3]); }
  template <class U>
  Eigen::Matrix<U, 4, 1> cast() const {
    return Eigen::Matrix<U, 4, 1>(static_cast<U>(p[0]), static_cast<U>(p[1]), static_cast<U>(p[2]), static_cast<U>(p[3]));
  }
};
struct alignas(16) PointXYZI {
  union {
    float data[4];
    struct {
      float x, y, z;
    };
  };
  union {
    struct {
      float intensity;
    };
    float data_c[4];
  };
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
  Vector4fMap getVector4fMap() { return Vector4fMap{data}; }
  Vector4fMapConst getVector4fMap() const { return Vector4fMapConst{data}; }
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");
}  // namespace pcl
#endif
