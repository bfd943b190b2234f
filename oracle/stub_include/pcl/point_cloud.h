// Stand-in for pcl/point_cloud.h (TEST INFRASTRUCTURE): the members of pcl::PointCloud the reference touches.
#ifndef DDLO_ORACLE_PCL_POINT_CLOUD_STUB
#define DDLO_ORACLE_PCL_POINT_CLOUD_STUB
#include <boost/shared_ptr.hpp>
#include <cstdint>
#include <vector>
#include <Eigen/Core>
namespace pcl {
template <class PointT>
class PointCloud {
 public:
  using Ptr = boost::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = boost::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT, Eigen::aligned_allocator<PointT>> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  PointT& at(std::size_t i) { return points.at(i); }
  void resize(std::size_t n) {
    points.resize(n);
    width = (std::uint32_t)n;
    height = 1;
  }
};
}  // namespace pcl
#endif
