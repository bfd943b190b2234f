// Convex and concave hulls of the keyframe positions, on the host (plain C++17, no CUDA).
//
// OdomNode::getSubmapKeyframes (src/odometry/odom.cc:1215-1297) adds to the submap the keyframes nearest to the
// current pose among the VERTICES of the convex hull and of the concave hull (alpha shape) of all keyframe positions
// (computeConvexHull :993-1028, computeConcaveHull :1030-1065).  The reference gets both from PCL 1.10
// (pcl::ConvexHull / pcl::ConcaveHull, surface/impl/convex_hull.hpp, concave_hull.hpp), which calls Qhull; neither is
// under /root/reference, so their published behaviour is restated here:
//
//   dimension        both classes look at the eigenvalues l0 <= l1 <= l2 of the covariance of the points and work in
//                    2-D when |l0| < eps or |l0 / l2| < 1e-3, else in 3-D (calculateInputDimension)
//   convex, 2-D      the points are projected onto a COORDINATE plane: with n = the normal of the triangle (first, last,
//                    middle point), xy unless n is within 10 degrees of the x or y axis, then yz / xz; hull vertices of
//                    the projected points ("qhull Tc": extreme points only)
//   convex, 3-D      vertices of the 3-D hull
//   concave, 2-D     the points are rotated into their principal plane; Delaunay triangulation ("qhull d QJ" = lower
//                    hull of the points lifted to z = x^2 + y^2); a triangle is GOOD when its circumradius is <= alpha;
//                    the hull consists of the edges of good triangles whose other side is not a good triangle
//   concave, 3-D     Delaunay tetrahedra; the hull consists of the triangles of circumradius <= alpha that do not
//                    separate two good tetrahedra.  Not implemented here (a 4-D hull): concave_hull_indices reports it
//                    and the keyframe store then selects from the k nearest keyframes and the convex hull only.
// Only the SET of hull vertices matters to the caller (it is sorted and made unique, odom.cc:1262-1264).
// Qhull decides near-degenerate configurations with its own tolerances and, for the concave hull, after a random
// joggle of the input (QJ); such configurations are outside what can be reproduced.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <map>
#include <utility>
#include <vector>

namespace ddlo {
namespace hull {

struct P3 {
  double x, y, z;
};

inline P3 sub(const P3& a, const P3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 cross(const P3& a, const P3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(const P3& a, const P3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// eigen-decomposition of a symmetric 3x3 (cyclic Jacobi); w ascending like pcl::eigen33, V[:, k] = k-th eigenvector
inline void eig_sym3(const double A[9], double w[3], double V[9]) {
  double a[9];
  for (int i = 0; i < 9; ++i) a[i] = A[i], V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = std::fabs(a[1]) + std::fabs(a[2]) + std::fabs(a[5]);
    if (off <= 1e-300 || off <= 1e-22 * (std::fabs(a[0]) + std::fabs(a[4]) + std::fabs(a[8]))) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[3 * p + q];
        if (apq == 0.0) continue;
        const double theta = (a[4 * q] - a[4 * p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const double akp = a[3 * k + p], akq = a[3 * k + q];
          a[3 * k + p] = c * akp - s * akq;
          a[3 * k + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const double apk = a[3 * p + k], aqk = a[3 * q + k];
          a[3 * p + k] = c * apk - s * aqk;
          a[3 * q + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = c * vkp - s * vkq;
          V[3 * k + q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int i, int j) { return a[4 * i] < a[4 * j]; });
  double Vs[9];
  for (int k = 0; k < 3; ++k) {
    w[k] = a[4 * order[k]];
    for (int r = 0; r < 3; ++r) Vs[3 * r + k] = V[3 * r + order[k]];
  }
  for (int i = 0; i < 9; ++i) V[i] = Vs[i];
}

inline void mean_and_covariance(const std::vector<P3>& p, P3& mean, double cov[9]) {
  mean = {0, 0, 0};
  for (const P3& q : p) mean.x += q.x, mean.y += q.y, mean.z += q.z;
  const double n = (double)std::max<size_t>(p.size(), 1);
  mean.x /= n, mean.y /= n, mean.z /= n;
  for (int i = 0; i < 9; ++i) cov[i] = 0.0;
  for (const P3& q : p) {
    const double d[3] = {q.x - mean.x, q.y - mean.y, q.z - mean.z};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov[3 * r + c] += d[r] * d[c] / n;
  }
}

// pcl::ConvexHull::calculateInputDimension / the same test in pcl::ConcaveHull::performReconstruction
inline int input_dimension(const std::vector<P3>& p, double* eigvecs = nullptr, P3* centroid = nullptr) {
  P3 mean;
  double cov[9], w[3], V[9];
  mean_and_covariance(p, mean, cov);
  eig_sym3(cov, w, V);
  if (eigvecs)
    for (int i = 0; i < 9; ++i) eigvecs[i] = V[i];
  if (centroid) *centroid = mean;
  if (std::fabs(w[0]) < std::numeric_limits<double>::epsilon() || std::fabs(w[0] / w[2]) < 1.0e-3) return 2;
  return 3;
}

// ---- 2-D convex hull: extreme points only (Andrew's monotone chain, strict turns) -------------------------------
inline std::vector<int> convex_hull_2d(const std::vector<std::array<double, 2>>& q) {
  const int n = (int)q.size();
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return q[a][0] < q[b][0] || (q[a][0] == q[b][0] && (q[a][1] < q[b][1] || (q[a][1] == q[b][1] && a < b))); });
  // duplicates of a position: only the first index can be a vertex
  idx.erase(std::unique(idx.begin(), idx.end(), [&](int a, int b) { return q[a][0] == q[b][0] && q[a][1] == q[b][1]; }), idx.end());
  const int m = (int)idx.size();
  if (m < 3) return idx;
  auto turn = [&](int a, int b, int c) { return (q[b][0] - q[a][0]) * (q[c][1] - q[a][1]) - (q[b][1] - q[a][1]) * (q[c][0] - q[a][0]); };
  std::vector<int> h(2 * m);
  int k = 0;
  for (int i = 0; i < m; ++i) {
    while (k >= 2 && turn(h[k - 2], h[k - 1], idx[i]) <= 0.0) --k;
    h[k++] = idx[i];
  }
  for (int i = m - 2, t = k + 1; i >= 0; --i) {
    while (k >= t && turn(h[k - 2], h[k - 1], idx[i]) <= 0.0) --k;
    h[k++] = idx[i];
  }
  h.resize(k - 1);
  return h;
}

// ---- 3-D convex hull (incremental, brute-force visibility; n is a few hundred at most) --------------------------
struct Hull3 {
  std::vector<std::array<int, 3>> faces;  // outward orientation (counter-clockwise seen from outside)
  bool ok = false;                        // false: the points are (numerically) coplanar
};

inline double orient(const P3& a, const P3& b, const P3& c, const P3& d) { return dot(cross(sub(b, a), sub(c, a)), sub(d, a)); }

inline Hull3 convex_hull_3d(const std::vector<P3>& p) {
  Hull3 H;
  const int n = (int)p.size();
  if (n < 4) return H;
  double scale = 0.0;
  for (const P3& q : p) scale = std::max({scale, std::fabs(q.x - p[0].x), std::fabs(q.y - p[0].y), std::fabs(q.z - p[0].z)});
  if (scale == 0.0) return H;
  const double eps = 1e-12 * scale * scale * scale;
  // a non-degenerate start: the farthest point from p0, the farthest from that line, the farthest from that plane
  int i0 = 0, i1 = -1, i2 = -1, i3 = -1;
  double best = 0.0;
  for (int i = 1; i < n; ++i) {
    const P3 d = sub(p[i], p[i0]);
    if (dot(d, d) > best) best = dot(d, d), i1 = i;
  }
  if (i1 < 0) return H;
  best = 0.0;
  for (int i = 0; i < n; ++i) {
    const P3 c = cross(sub(p[i1], p[i0]), sub(p[i], p[i0]));
    if (dot(c, c) > best) best = dot(c, c), i2 = i;
  }
  if (i2 < 0 || best <= 1e-24 * scale * scale * scale * scale) return H;
  best = 0.0;
  for (int i = 0; i < n; ++i) {
    const double v = std::fabs(orient(p[i0], p[i1], p[i2], p[i]));
    if (v > best) best = v, i3 = i;
  }
  if (i3 < 0 || best <= eps) return H;
  if (orient(p[i0], p[i1], p[i2], p[i3]) > 0) std::swap(i1, i2);  // now i3 lies below (i0, i1, i2): that face looks away from it
  std::vector<std::array<int, 3>> F = {{i0, i1, i2}, {i0, i3, i1}, {i1, i3, i2}, {i2, i3, i0}};
  std::vector<char> used(n, 0);
  used[i0] = used[i1] = used[i2] = used[i3] = 1;
  for (int i = 0; i < n; ++i) {
    if (used[i]) continue;
    std::vector<char> vis(F.size(), 0);
    bool any = false;
    for (size_t f = 0; f < F.size(); ++f)
      if (orient(p[F[f][0]], p[F[f][1]], p[F[f][2]], p[i]) > eps) vis[f] = 1, any = true;
    if (!any) continue;  // inside or on the hull
    // horizon: directed edges of visible faces whose reverse belongs to a face that stays
    std::map<std::pair<int, int>, int> edge_face;
    for (size_t f = 0; f < F.size(); ++f)
      for (int e = 0; e < 3; ++e) edge_face[{F[f][e], F[f][(e + 1) % 3]}] = (int)f;
    std::vector<std::array<int, 3>> G;
    for (size_t f = 0; f < F.size(); ++f)
      if (!vis[f]) G.push_back(F[f]);
    for (size_t f = 0; f < F.size(); ++f) {
      if (!vis[f]) continue;
      for (int e = 0; e < 3; ++e) {
        const int a = F[f][e], b = F[f][(e + 1) % 3];
        const auto it = edge_face.find({b, a});
        if (it != edge_face.end() && !vis[it->second]) G.push_back({a, b, i});
      }
    }
    F.swap(G);
  }
  H.faces = F;
  H.ok = true;
  return H;
}

// ---- what the callers ask for -----------------------------------------------------------------------------------
// pcl::ConvexHull::reconstruct + getHullPointIndices as a sorted set of indices into `p`
inline std::vector<int> convex_hull_indices(const std::vector<P3>& p) {
  std::vector<int> out;
  const int n = (int)p.size();
  if (n < 3) return out;
  if (input_dimension(p) == 3) {
    const Hull3 H = convex_hull_3d(p);
    if (H.ok) {
      for (const auto& f : H.faces) out.insert(out.end(), f.begin(), f.end());
      std::sort(out.begin(), out.end());
      out.erase(std::unique(out.begin(), out.end()), out.end());
      return out;
    }
  }
  // 2-D: which coordinate plane (performReconstruction2D)
  const P3 p0 = p[0], p1 = p[n - 1], p2 = p[n / 2];
  P3 nrm = cross(sub(p1, p0), sub(p2, p0));
  const double len = std::sqrt(dot(nrm, nrm));
  bool xy = true, yz = true, xz = true;
  if (len > 0.0) {
    const double thresh = std::cos(0.174532925);
    const double tx = std::fabs(nrm.x / len), ty = std::fabs(nrm.y / len), tz = std::fabs(nrm.z / len);
    if (tz > thresh) xz = false, yz = false;
    if (tx > thresh) xz = false, xy = false;
    if (ty > thresh) xy = false, yz = false;
  }
  std::vector<std::array<double, 2>> q(n);
  for (int i = 0; i < n; ++i) {
    if (xy)
      q[i] = {p[i].x, p[i].y};
    else if (yz)
      q[i] = {p[i].y, p[i].z};
    else if (xz)
      q[i] = {p[i].x, p[i].z};
    else
      q[i] = {p[i].x, p[i].y};
  }
  out = convex_hull_2d(q);
  std::sort(out.begin(), out.end());
  return out;
}

// pcl::ConcaveHull::reconstruct + getHullPointIndices with setAlpha(alpha) as a sorted set of indices into `p`.
// *dimension receives 2 or 3; in the 3-D case nothing is computed and the result is empty (see the header comment).
inline std::vector<int> concave_hull_indices(const std::vector<P3>& p, double alpha, int* dimension = nullptr) {
  std::vector<int> out;
  const int n = (int)p.size();
  double V[9];
  P3 c;
  const int dim = n >= 3 ? input_dimension(p, V, &c) : 2;
  if (dimension) *dimension = dim;
  if (n < 3 || dim == 3) return out;
  // rotate into the principal plane: x along the largest, y along the middle eigenvector (transform1 of PCL)
  std::vector<P3> lifted(n);
  for (int i = 0; i < n; ++i) {
    const P3 d = sub(p[i], c);
    const double u = d.x * V[2] + d.y * V[5] + d.z * V[8];
    const double v = d.x * V[1] + d.y * V[4] + d.z * V[7];
    lifted[i] = {u, v, u * u + v * v};
  }
  const Hull3 H = convex_hull_3d(lifted);  // its lower faces are the Delaunay triangles
  if (!H.ok) return out;                   // all points on one circle or line: no triangulation
  const int nf = (int)H.faces.size();
  std::vector<char> lower(nf, 0), good(nf, 0);
  for (int f = 0; f < nf; ++f) {
    const P3 &a = lifted[H.faces[f][0]], &b = lifted[H.faces[f][1]], &d = lifted[H.faces[f][2]];
    const P3 nrm = cross(sub(b, a), sub(d, a));
    lower[f] = nrm.z < 0.0;  // !facet->upperdelaunay
    if (!lower[f]) continue;
    // circumradius of the 2-D triangle (distance of a vertex from the Voronoi centre)
    const double ax = a.x, ay = a.y, bx = b.x - ax, by = b.y - ay, dx = d.x - ax, dy = d.y - ay;
    const double den = 2.0 * (bx * dy - by * dx);
    if (den == 0.0) continue;
    const double ux = (dy * (bx * bx + by * by) - by * (dx * dx + dy * dy)) / den;
    const double uy = (bx * (dx * dx + dy * dy) - dx * (bx * bx + by * by)) / den;
    good[f] = std::sqrt(ux * ux + uy * uy) <= alpha;
  }
  std::map<std::pair<int, int>, int> edge_face;
  for (int f = 0; f < nf; ++f)
    for (int e = 0; e < 3; ++e) edge_face[{H.faces[f][e], H.faces[f][(e + 1) % 3]}] = f;
  std::vector<char> on(n, 0);
  for (int f = 0; f < nf; ++f) {
    if (!good[f]) continue;
    for (int e = 0; e < 3; ++e) {
      const int a = H.faces[f][e], b = H.faces[f][(e + 1) % 3];
      const auto it = edge_face.find({b, a});
      const int g = it == edge_face.end() ? -1 : it->second;
      if (g < 0 || !lower[g] || !good[g]) on[a] = on[b] = 1;  // the ridge's other side is upper-Delaunay or not good
    }
  }
  for (int i = 0; i < n; ++i)
    if (on[i]) out.push_back(i);
  return out;
}

}  // namespace hull
}  // namespace ddlo
