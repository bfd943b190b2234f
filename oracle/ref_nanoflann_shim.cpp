// ORACLE / TEST INFRASTRUCTURE ONLY — never linked into, loaded by, or called from the product path.
//
// Thin C wrapper that compiles the reference's OWN vendored nanoflann 1.3.2
//   /root/reference/dynamic_direct_lidar_odometry/include/nano_gicp/impl/nanoflann_impl.hpp
// where it lies (include path given by oracle/Makefile; no reference source is copied into this
// repo) into oracle/_ref/libnanoflann_ref.so.  It reproduces exactly how the reference drives it:
//
//   * index type  KDTreeSingleIndexAdaptor<SO3_Adaptor<float, Adaptor>, Adaptor, 3, int>
//                                                        (nano_gicp/nanoflann.hpp:107-109)
//   * leaf_max_size = 100                                (nano_gicp/nanoflann.hpp:119)
//   * setInputCloud -> buildIndex()                      (nano_gicp/nanoflann.hpp:137-143)
//   * nearestKSearch: KNNResultSet<float,int>(k), init(), findNeighbors(rs, q, SearchParams())
//                                                        (nano_gicp/nanoflann.hpp:146-156)
//   * adaptor: kdtree_get_pt(idx, dim) returns x/y/z, kdtree_get_bbox -> false
//                                                        (nano_gicp/nanoflann.hpp:182-203)
//
// The PCL point-cloud adaptor of the reference (nanoflann.hpp needs <pcl/...>, absent here) is
// replaced by the POD adaptor below; everything under it is the reference's code, unmodified.
#include <cstddef>
#include <cstdint>
#include <vector>

#include <nano_gicp/impl/nanoflann_impl.hpp>

namespace {

struct PodAdaptor {
  const float* xyz = nullptr;  // n * stride floats
  size_t n = 0;
  size_t stride = 4;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, int dim) const {
    if (dim >= 0 && dim < 3) return xyz[idx * stride + dim];
    return 0.0f;
  }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const { return false; }
};

using RefTree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::SO3_Adaptor<float, PodAdaptor>, PodAdaptor, 3, int>;

struct RefIndex {
  std::vector<float> pts;
  PodAdaptor adaptor;
  RefTree tree;
  RefIndex() : tree(3, adaptor, nanoflann::KDTreeSingleIndexAdaptorParams(100)) {}
};

}  // namespace

extern "C" {

void* ref_kdtree_build(const float* xyz, int n, int stride_floats) {
  RefIndex* ix = new RefIndex();
  ix->pts.assign(xyz, xyz + static_cast<size_t>(n) * stride_floats);
  ix->adaptor.xyz = ix->pts.data();
  ix->adaptor.n = static_cast<size_t>(n);
  ix->adaptor.stride = static_cast<size_t>(stride_floats);
  ix->tree.buildIndex();
  return ix;
}

void ref_kdtree_free(void* h) { delete static_cast<RefIndex*>(h); }

// query: 3 floats. Returns the number of neighbours found (== k unless n < k).
int ref_kdtree_knn(void* h, const float* query, int k, int* idx_out, float* d2_out) {
  RefIndex* ix = static_cast<RefIndex*>(h);
  nanoflann::KNNResultSet<float, int> rs(k);
  rs.init(idx_out, d2_out);
  ix->tree.findNeighbors(rs, query, nanoflann::SearchParams());
  return static_cast<int>(rs.size());
}

}  // extern "C"
