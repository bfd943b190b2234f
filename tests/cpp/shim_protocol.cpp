// C++ host-side test of the NanoGICP shim: replays the call sequence of OdomNode
// (odom.cc:480-532 initializeInputTarget / setInputSources, :745-793 scanMatching) on scans read from a
// file, and prints every transform so that the Python harness can compare them with the same
// protocol driven through the C ABI directly.
//
//   shim_protocol scans.bin      scans.bin: int32 count, then per scan int32 n and n*4 float32 (x y z 1)
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <nano_gicp/nano_gicp.hpp>

using PointType = ddlo_shim::PointXYZI;
using CloudT = ddlo_shim::Cloud<PointType>;

static std::vector<CloudT::Ptr> load(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    std::perror(path);
    std::exit(2);
  }
  int count = 0;
  if (std::fread(&count, 4, 1, f) != 1) std::exit(2);
  std::vector<CloudT::Ptr> scans;
  for (int s = 0; s < count; ++s) {
    int n = 0;
    if (std::fread(&n, 4, 1, f) != 1) std::exit(2);
    std::vector<float> xyzw(4 * (size_t)n);
    if (std::fread(xyzw.data(), 4, xyzw.size(), f) != xyzw.size()) std::exit(2);
    auto c = std::make_shared<CloudT>();
    c->points.resize(n);
    for (int i = 0; i < n; ++i) {
      c->points[i].x = xyzw[4 * i];
      c->points[i].y = xyzw[4 * i + 1];
      c->points[i].z = xyzw[4 * i + 2];
    }
    scans.push_back(c);
  }
  std::fclose(f);
  return scans;
}

static void print_T(const char* tag, int frame, const ddlo_shim::Matrix4f& T, bool conv, int iters) {
  std::printf("%s %d %d %d", tag, frame, conv ? 1 : 0, iters);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) std::printf(" %.9g", T(r, c));
  std::printf("\n");
}

static ddlo_shim::Matrix4f mul(const ddlo_shim::Matrix4f& a, const ddlo_shim::Matrix4f& b) {
  ddlo_shim::Matrix4f r = ddlo_shim::Matrix4f::Zero();
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
    nano_gicp::NanoGICP<PointType, PointType> gicp_s2s_, gicp_s2m_;
    // OdomNode constructor (odom.cc:92-112)
    for (auto* g : {&gicp_s2s_, &gicp_s2m_}) {
      g->setCorrespondenceRandomness(10);
      g->setMaxCorrespondenceDistance(1.0);
      g->setMaximumIterations(32);
      g->setTransformationEpsilon(0.01);
      g->setEuclideanFitnessEpsilon(0.01);
      g->setRANSACIterations(5);
      g->setRANSACOutlierRejectionThreshold(1.0);
      std::shared_ptr<int> none;
      g->setSearchMethodSource(none, true);
      g->setSearchMethodTarget(none, true);
    }
    // initializeInputTarget (odom.cc:480-516): first scan is the S2S target and the first keyframe
    CloudT::Ptr target_cloud_ = scans[0];
    gicp_s2s_.setInputTarget(target_cloud_);
    gicp_s2s_.calculateTargetCovariances();
    gicp_s2s_.setInputSource(target_cloud_);  // keyframe_cloud_ (identity pose)
    gicp_s2s_.calculateSourceCovariances();
    nano_gicp::DeviceCovariances keyframe_normals = gicp_s2s_.getSourceCovariances();
    gicp_s2m_.setInputTarget(target_cloud_);  // submap = first keyframe
    gicp_s2m_.setTargetCovariances(keyframe_normals);

    ddlo_shim::Matrix4f T_ = ddlo_shim::Matrix4f::Identity();
    for (size_t f = 1; f < scans.size(); ++f) {
      CloudT::Ptr registration_scan_ = scans[f];
      // setInputSources (odom.cc:518-532)
      gicp_s2s_.setInputSource(registration_scan_);
      gicp_s2m_.registerInputSource(registration_scan_);
      gicp_s2m_.source_kdtree_ = gicp_s2s_.source_kdtree_;
      gicp_s2m_.source_covs_.clear();
      // scanMatching (odom.cc:745-793)
      CloudT aligned;
      gicp_s2s_.align(aligned);
      ddlo_shim::Matrix4f T_S2S = gicp_s2s_.getFinalTransformation();
      print_T("s2s", (int)f, T_S2S, gicp_s2s_.hasConverged(), gicp_s2s_.getNrIterations());
      ddlo_shim::Matrix4f T_s2s_ = mul(T_, T_S2S);  // propagateS2S
      gicp_s2m_.source_covs_ = gicp_s2s_.source_covs_;
      gicp_s2s_.swapSourceAndTarget();
      gicp_s2m_.align(aligned, T_s2s_);
      T_ = gicp_s2m_.getFinalTransformation();
      print_T("s2m", (int)f, T_, gicp_s2m_.hasConverged(), gicp_s2m_.getNrIterations());
      std::vector<double> residuals;
      gicp_s2m_.getResiduals(residuals, T_);
      double sum = 0.0;
      for (double r : residuals) sum += r;
      std::printf("res %d %zu %.12g %zu\n", (int)f, residuals.size(), sum, aligned.size());
    }
    // host round trip of the covariance vector type
    ddlo_shim::Matrix4dVector host = keyframe_normals;
    nano_gicp::DeviceCovariances back;
    back = host;
    std::printf("covs %zu %zu %.17g\n", host.size(), back.size(), host.empty() ? 0.0 : host[0](0, 0));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
