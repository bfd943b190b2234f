// The NanoGICP shim in its PRODUCTION configuration: compiled with <pcl/point_cloud.h>, <pcl/point_types.h>,
// <Eigen/Core> and <Eigen/StdVector> on the include path (here: the stand-in headers of oracle/stub_include, a
// test-only include path - the real PCL 1.10 / Eigen 3.3 are not installed in this image), so that
// include/nano_gicp/nano_gicp.hpp takes the branch OdomNode would take:
//   nano_gicp::NanoGICP<pcl::PointXYZI, pcl::PointXYZI>        (ddlo.h:86-90, odom.h:159-160)
//   pcl::PointCloud<pcl::PointXYZI>::Ptr clouds, Eigen::Matrix4f transforms,
//   std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> covariance vectors.
// It replays OdomNode's calls - initializeInputTarget (odom.cc:480-516), setInputSources (:518-532), scanMatching
// (:745-793), the keyframe update (:1140-1149) and the submap hand-over (:780-784) - and prints the same lines as
// shim_protocol.cpp (the POD branch); the Python harness requires the two outputs to be identical bit for bit.
//
//   shim_pcl_branch scans.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Eigen/Core>
#include <Eigen/StdVector>

#include <nano_gicp/nano_gicp.hpp>

#if !defined(DDLO_HAVE_PCL) || !defined(DDLO_HAVE_EIGEN)
#error "this test must be compiled with PCL and Eigen headers (or their stand-ins) on the include path"
#endif

using PointType = pcl::PointXYZI;
using CovVector = std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>;

static std::vector<pcl::PointCloud<PointType>::Ptr> load(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    std::perror(path);
    std::exit(2);
  }
  int count = 0;
  if (std::fread(&count, 4, 1, f) != 1) std::exit(2);
  std::vector<pcl::PointCloud<PointType>::Ptr> scans;
  for (int s = 0; s < count; ++s) {
    int n = 0;
    if (std::fread(&n, 4, 1, f) != 1) std::exit(2);
    std::vector<float> xyzw(4 * (size_t)n);
    if (std::fread(xyzw.data(), 4, xyzw.size(), f) != xyzw.size()) std::exit(2);
    pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>);
    c->resize(n);
    for (int i = 0; i < n; ++i) {
      PointType& p = c->points[i];
      p.x = xyzw[4 * i], p.y = xyzw[4 * i + 1], p.z = xyzw[4 * i + 2];
      p.data[3] = 1.0f;
      p.intensity = 0.0f;
    }
    scans.push_back(c);
  }
  std::fclose(f);
  return scans;
}

static void print_T(const char* tag, int frame, const Eigen::Matrix4f& T, bool conv, int iters) {
  std::printf("%s %d %d %d", tag, frame, conv ? 1 : 0, iters);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) std::printf(" %.9g", T(r, c));
  std::printf("\n");
}

// the float product with the summation order of shim_protocol.cpp (the two programs must print identical numbers)
static Eigen::Matrix4f mul(const Eigen::Matrix4f& a, const Eigen::Matrix4f& b) {
  Eigen::Matrix4f r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s scans.bin\n", argv[0]);
    return 2;
  }
  try {
    auto scans = load(argv[1]);
    nano_gicp::NanoGICP<PointType, PointType> gicp_s2s_, gicp_s2m_;  // odom.h:159-160
    // OdomNode constructor (odom.cc:92-112)
    for (auto* g : {&gicp_s2s_, &gicp_s2m_}) {
      g->setCorrespondenceRandomness(10);
      g->setMaxCorrespondenceDistance(1.0);
      g->setMaximumIterations(32);
      g->setTransformationEpsilon(0.01);
      g->setEuclideanFitnessEpsilon(0.01);
      g->setRANSACIterations(5);
      g->setRANSACOutlierRejectionThreshold(1.0);
      pcl::PointCloud<PointType>::Ptr none;  // pcl::search::KdTree<PointType>::Ptr in the reference: ignored either way
      g->setSearchMethodSource(none, true);
      g->setSearchMethodTarget(none, true);
    }
    std::vector<CovVector> keyframe_normals_;  // odom.h:107-108
    CovVector submap_normals_;                 // odom.h:131

    // initializeInputTarget (odom.cc:480-516)
    pcl::PointCloud<PointType>::Ptr target_cloud_ = scans[0];
    gicp_s2s_.setInputTarget(target_cloud_);
    gicp_s2s_.calculateTargetCovariances();
    pcl::PointCloud<PointType>::Ptr keyframe_cloud_ = target_cloud_;  // identity pose
    gicp_s2s_.setInputSource(keyframe_cloud_);
    gicp_s2s_.calculateSourceCovariances();
    keyframe_normals_.push_back(gicp_s2s_.getSourceCovariances());  // device vector -> std::vector<Matrix4d, aligned_allocator>
    // getSubmapKeyframes (odom.cc:1298-1313): the submap is the first keyframe, its normals are concatenated on the host
    pcl::PointCloud<PointType>::Ptr submap_cloud_ = keyframe_cloud_;
    submap_normals_.insert(submap_normals_.end(), keyframe_normals_[0].begin(), keyframe_normals_[0].end());
    gicp_s2m_.setInputTarget(submap_cloud_);             // odom.cc:780
    gicp_s2m_.setTargetCovariances(submap_normals_);     // odom.cc:783: host vector, uploaded

    Eigen::Matrix4f T_ = Eigen::Matrix4f::Identity();
    for (size_t f = 1; f < scans.size(); ++f) {
      pcl::PointCloud<PointType>::Ptr registration_scan_ = scans[f];
      // setInputSources (odom.cc:518-532)
      gicp_s2s_.setInputSource(registration_scan_);
      gicp_s2m_.registerInputSource(registration_scan_);
      gicp_s2m_.source_kdtree_ = gicp_s2s_.source_kdtree_;
      gicp_s2m_.source_covs_.clear();
      // scanMatching (odom.cc:745-793)
      pcl::PointCloud<PointType>::Ptr aligned(new pcl::PointCloud<PointType>);
      gicp_s2s_.align(*aligned);
      Eigen::Matrix4f T_S2S = gicp_s2s_.getFinalTransformation();
      print_T("s2s", (int)f, T_S2S, gicp_s2s_.hasConverged(), gicp_s2s_.getNrIterations());
      Eigen::Matrix4f T_s2s_ = mul(T_, T_S2S);  // propagateS2S
      gicp_s2m_.source_covs_ = gicp_s2s_.source_covs_;
      gicp_s2s_.swapSourceAndTarget();
      gicp_s2m_.align(*aligned, T_s2s_);
      T_ = gicp_s2m_.getFinalTransformation();
      print_T("s2m", (int)f, T_, gicp_s2m_.hasConverged(), gicp_s2m_.getNrIterations());
      std::vector<double> residuals;
      gicp_s2m_.getResiduals(residuals, T_);
      double sum = 0.0;
      for (double r : residuals) sum += r;
      std::printf("res %d %zu %.12g %zu\n", (int)f, residuals.size(), sum, aligned->size());
    }
    // host round trip of the covariance vector type
    CovVector host = keyframe_normals_[0];
    nano_gicp::DeviceCovariances back;
    back = host;
    std::printf("covs %zu %zu %.17g\n", host.size(), back.size(), host.empty() ? 0.0 : host[0](0, 0));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
