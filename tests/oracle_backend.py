"""Backend of dynamic_direct_lidar_odometry_b200.odometry_loop for the CPU oracle (test infrastructure only)."""
import numpy as np


class OracleBackend:
    """CPU restatement (oracle/), test infrastructure only."""

    name = "oracle"

    def __init__(self, pyoracle):
        self.po = pyoracle

    def cloud(self, points):
        return self.po.Cloud(points)

    def voxel_filter(self, cloud, leaf):
        return self.po.Cloud(self.po.voxel_filter(cloud.points, leaf))

    def size(self, cloud):
        return cloud.n

    def transform(self, cloud, T):
        # float arithmetic of pcl::transformPointCloud as Eigen evaluates it: (r0 x + r1 y) + (r2 z + t)
        T = np.asarray(T, dtype=np.float32)
        p = cloud.points[:, :3].astype(np.float32)
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        out = np.empty((len(p), 4), dtype=np.float32)
        for r in range(3):
            out[:, r] = (T[r, 0] * x + T[r, 1] * y) + (T[r, 2] * z + T[r, 3])
        out[:, 3] = 1.0
        return self.po.Cloud(out)

    def concat_clouds(self, parts):
        return self.po.Cloud(np.concatenate([c.points for c in parts], axis=0))

    def concat_covs(self, parts):
        return np.concatenate(parts, axis=0)

    def engine(self):
        return self.po.NanoGICP()

    def share_source(self, s2m, s2s):
        s2m.clearSourceCovariances()  # the tree is rebuilt by the oracle's registerInputSource path when needed

    def hand_over_source_covs(self, s2m, s2s):
        s2m.setSourceCovariances(s2s.getSourceCovariances())

    def hull_indices(self, positions, alpha):
        """pcl::ConvexHull / pcl::ConcaveHull vertex sets through Qhull itself (scipy), independent of csrc/hull.hpp;
        only the 2-D case (positions within PCL's planarity test), which is what the test sequences produce"""
        return qhull_convex(positions), qhull_concave(positions, alpha)

    def sync(self):
        pass


def pcl_dimension(positions) -> int:
    """calculateInputDimension of pcl::ConvexHull / the same test in pcl::ConcaveHull"""
    p = np.asarray(positions, dtype=np.float64)
    w = np.linalg.eigvalsh(np.cov((p - p.mean(0)).T, bias=True))
    return 2 if abs(w[0]) < np.finfo(np.float64).eps or abs(w[0] / w[2]) < 1e-3 else 3


def qhull_convex(positions):
    from scipy.spatial import ConvexHull

    p = np.asarray(positions, dtype=np.float64)
    if pcl_dimension(p) == 3:
        return sorted(int(v) for v in ConvexHull(p).vertices)
    # performReconstruction2D: the coordinate plane is chosen from the normal of (first, last, middle point)
    nrm = np.cross(p[-1] - p[0], p[len(p) // 2] - p[0])
    ln = np.linalg.norm(nrm)
    xy = yz = xz = True
    if ln > 0:
        tx, ty, tz = np.abs(nrm / ln)
        th = np.cos(0.174532925)
        if tz > th:
            xz = yz = False
        if tx > th:
            xz = xy = False
        if ty > th:
            xy = yz = False
    cols = [0, 1] if xy else ([1, 2] if yz else ([0, 2] if xz else [0, 1]))
    return sorted(int(v) for v in ConvexHull(p[:, cols]).vertices)


def qhull_concave(positions, alpha):
    from scipy.spatial import Delaunay

    p = np.asarray(positions, dtype=np.float64)
    if pcl_dimension(p) == 3:
        return []
    c = p.mean(0)
    _, V = np.linalg.eigh(np.cov((p - c).T, bias=True))
    uv = np.stack([(p - c) @ V[:, 2], (p - c) @ V[:, 1]], axis=1)  # transform1: largest eigenvector -> x, middle -> y
    tri = Delaunay(uv)

    def circumradius(t):
        a, b, d = uv[t]
        bx, by = b - a
        dx, dy = d - a
        den = 2 * (bx * dy - by * dx)
        ux = (dy * (bx * bx + by * by) - by * (dx * dx + dy * dy)) / den
        uy = (bx * (dx * dx + dy * dy) - dx * (bx * bx + by * by)) / den
        return float(np.hypot(ux, uy))

    good = [circumradius(t) <= alpha for t in tri.simplices]
    on = set()
    for f, t in enumerate(tri.simplices):
        if not good[f]:
            continue
        for e in range(3):
            nb = tri.neighbors[f][e]  # the triangle opposite vertex e, -1 outside the triangulation
            if nb < 0 or not good[nb]:
                on.update((int(t[(e + 1) % 3]), int(t[(e + 2) % 3])))
    return sorted(on)
