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

    def sync(self):
        pass
