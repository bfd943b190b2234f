"""The caller side of the registration path, as OdomNode drives it (harness for config C3).

This replays, around two NanoGICP engines, exactly the calls of
    OdomNode::initializeInputTarget   odom.cc:480-516
    OdomNode::setInputSources         odom.cc:518-532
    OdomNode::scanMatching            odom.cc:745-793
    OdomNode::propagateS2S / S2M      odom.cc:921-955
    OdomNode::updateKeyframes         odom.cc:1067-1150
    OdomNode::getSubmapKeyframes      odom.cc:1215-1315 (nearest keyframes + nearest convex- / concave-hull keyframes, concatenation)
so that the sequence benchmark (benchmarks/c3_sequence.py) and the sequence parity test exercise the
engine with the reference's own protocol: S2S align, covariance hand-over to S2M, swapSourceAndTarget,
shared source index, keyframe covariances computed through the S2S source slot, submap clouds and
covariances concatenated per selected keyframe and injected with setInputTarget/setTargetCovariances
only when the selection changed.

It is written against a tiny backend interface (`GpuBackend`: every cloud, covariance vector and submap
stays on the device) so that the tests can replay the very same loop on their CPU checker with a
backend of their own (tests/oracle_backend.py); this package has no CPU path.  The hull part of the keyframe selection
(submap_kcv / submap_kcc > 0) asks the backend for the hull vertices: the GPU backend uses the C++ hull code of the
keyframe store (csrc/hull.hpp), the tests' backend uses Qhull through scipy.  The same loop written in C++ on the
keyframe store of the C ABI (ddlo_keyframes_*) is tests/cpp/odometry_sequence.cpp.  Out of scope, as in SURVEY.md §8:
IMU prior, ROS I/O.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


# ------------------------------------------------------------------------------------------- backends
class GpuBackend:
    """Clouds / covariances are device handles of the C ABI; nothing crosses PCIe except scans in and poses out."""

    name = "gpu"

    def __init__(self, rt):
        from . import nano_gicp as ng

        self.ng, self.rt = ng, rt

    def cloud(self, points):
        return self.ng.PointCloud(self.rt, points)

    def voxel_filter(self, cloud, leaf):  # vf_scan_.filter / vf_submap_.filter (odom.cc:469-474, 1133-1137), on the device
        return cloud.voxel_filtered(leaf)

    def size(self, cloud):
        return cloud.size()

    def transform(self, cloud, T):  # pcl::transformPointCloud
        return cloud.transformed(np.asarray(T, dtype=np.float32))

    def concat_clouds(self, parts):  # *submap_cloud += *keyframe
        return self.ng.PointCloud.concat(self.rt, parts)

    def concat_covs(self, parts):  # submap_normals_.insert(...)
        return self.ng.Covariances.concat(self.rt, parts)

    def engine(self):
        return self.ng.NanoGICP(self.rt)

    def share_source(self, s2m, s2s):  # odom.cc:527-531
        s2m.source_kdtree_ = s2s.source_kdtree_
        s2m.source_covs_ = None

    def hand_over_source_covs(self, s2m, s2s):  # odom.cc:765
        s2m.source_covs_ = s2s.source_covs_

    def hull_indices(self, positions, alpha):  # computeConvexHull / computeConcaveHull (odom.cc:993-1065) -> csrc/hull.hpp
        cc = self.ng.hull_concave(positions, alpha)
        return self.ng.hull_convex(positions), ([] if cc is None else cc)

    def sync(self):
        self.rt.synchronize()


# ------------------------------------------------------------------------------------------- the loop
@dataclass
class LoopConfig:
    k_correspondences_s2s: int = 20   # gicp_s2s kCorrespondences (engine default; the DLO yaml uses 10)
    k_correspondences_s2m: int = 20
    max_correspondence_distance: Optional[float] = None  # None = engine default (FLT_MAX)
    max_iterations: int = 64
    keyframe_thresh_dist: float = 1.0   # metres (odom.cc:1169, "rebuilt every 1 m" in SURVEY.md §8d C3)
    keyframe_thresh_rot: float = 15.0   # degrees
    submap_knn: int = 10                # odomNode/submap/keyframe/knn
    submap_kcv: int = 0                 # .../kcv: nearest keyframes among the convex-hull vertices (the reference's default is 10; 0 = off)
    submap_kcc: int = 0                 # .../kcc: the same for the concave hull (alpha = keyframe_thresh_dist, odom.cc:1175)
    voxel_leaf_scan: Optional[float] = None    # vf_scan_ leaf size (preprocessPoints, odom.cc:469-474); None = off
    voxel_leaf_submap: Optional[float] = None  # vf_submap_ leaf size applied to every new keyframe (odom.cc:494-499, 1133-1137)


@dataclass
class FrameRecord:
    T_s2s: np.ndarray
    T: np.ndarray
    s2s_iterations: int
    s2m_iterations: int
    s2s_converged: bool
    s2m_converged: bool
    new_keyframe: bool
    submap_changed: bool
    submap_points: int
    seconds: float
    residual_mean: float


@dataclass
class Keyframe:
    position: np.ndarray  # float32 xyz
    rotation: np.ndarray  # 3x3 float32
    cloud: object         # world-frame cloud (backend handle)
    covs: object          # its covariances (backend handle)
    n: int


def _rot_angle_deg(Ra: np.ndarray, Rb: np.ndarray) -> float:
    """angle of Ra * Rb^-1 in degrees (the quaternion formula of odom.cc:1105-1108, via the trace)"""
    c = (np.trace(Ra.astype(np.float64) @ Rb.astype(np.float64).T) - 1.0) / 2.0
    return math.degrees(math.acos(min(1.0, max(-1.0, c))))


class OdometryLoop:
    def __init__(self, backend, cfg: LoopConfig = LoopConfig()):
        self.b, self.cfg = backend, cfg
        self.s2s, self.s2m = backend.engine(), backend.engine()
        for e, k in ((self.s2s, cfg.k_correspondences_s2s), (self.s2m, cfg.k_correspondences_s2m)):
            e.setCorrespondenceRandomness(k)
            e.setMaximumIterations(cfg.max_iterations)
            if cfg.max_correspondence_distance is not None:
                e.setMaxCorrespondenceDistance(cfg.max_correspondence_distance)
        self.T = np.eye(4, dtype=np.float32)
        self.T_s2s_prev = np.eye(4, dtype=np.float32)
        self.keyframes: List[Keyframe] = []
        self.submap_idx_prev: List[int] = []
        self.records: List[FrameRecord] = []
        self._convex: List[int] = []   # keyframe_convex_ / keyframe_concave_
        self._concave: List[int] = []
        self._initialised = False

    # odom.cc:480-516
    def _initialize_input_target(self, scan_cloud, n):
        self.s2s.setInputTarget(scan_cloud)
        self.s2s.calculateTargetCovariances()
        first = self.b.transform(scan_cloud, self.T)
        if self.cfg.voxel_leaf_submap:
            first = self.b.voxel_filter(first, self.cfg.voxel_leaf_submap)
        self.s2s.setInputSource(first)  # temporary storage, overwritten by the next setInputSources()
        self.s2s.calculateSourceCovariances()
        self.keyframes.append(Keyframe(self.T[:3, 3].copy(), self.T[:3, :3].copy(), first, self.s2s.getSourceCovariances(), self.b.size(first)))
        self._initialised = True

    @staticmethod
    def _push_submap_indices(dists, k, frames, out):  # odom.cc:1178-1213: every frame at most as far as the k-th nearest
        if not dists or k <= 0:
            return
        kth = sorted(dists)[min(k, len(dists)) - 1]
        out.extend(f for f, v in zip(frames, dists) if v <= kth)

    # odom.cc:1215-1315
    def _submap_selection(self, position) -> List[int]:
        pos = position.astype(np.float32)
        # float differences, double squares and root, rounded to float (sqrt(pow(float, 2) + ...) assigned to a float)
        d = [float(np.float32(np.sqrt(np.sum((pos - k.position).astype(np.float64) ** 2)))) for k in self.keyframes]
        sel: List[int] = []
        self._push_submap_indices(d, self.cfg.submap_knn, list(range(len(d))), sel)
        if self.cfg.submap_kcv > 0 or self.cfg.submap_kcc > 0:
            n = len(self.keyframes)
            pts = np.array([k.position for k in self.keyframes], dtype=np.float64)
            convex, concave = self.b.hull_indices(pts, self.cfg.keyframe_thresh_dist) if n >= 4 else ([], [])
            if n >= 4:  # computeConvexHull: at least 4 keyframes, else the previous (empty) list stays
                self._convex = convex
            if n >= 5:  # computeConcaveHull: at least 5
                self._concave = concave
            self._push_submap_indices([d[i] for i in self._convex], self.cfg.submap_kcv, self._convex, sel)
            self._push_submap_indices([d[i] for i in self._concave], self.cfg.submap_kcc, self._concave, sel)
        return sorted(set(sel))

    # odom.cc:1067-1150
    def _update_keyframes(self, scan_cloud, n) -> bool:
        pos, rot = self.T[:3, 3], self.T[:3, :3]
        d = [float(np.float32(np.sqrt(np.sum((pos - k.position).astype(np.float64) ** 2)))) for k in self.keyframes]
        closest = int(np.argmin(d))
        num_nearby = sum(1 for v in d if v <= self.cfg.keyframe_thresh_dist * 1.5)
        dd, theta = d[closest], _rot_angle_deg(rot, self.keyframes[closest].rotation)
        new = dd > self.cfg.keyframe_thresh_dist or theta > self.cfg.keyframe_thresh_rot
        if dd <= self.cfg.keyframe_thresh_dist:
            new = False
        if dd <= self.cfg.keyframe_thresh_dist and theta > self.cfg.keyframe_thresh_rot and num_nearby <= 1:
            new = True
        if new:
            kf = self.b.transform(scan_cloud, self.T)
            if self.cfg.voxel_leaf_submap:
                kf = self.b.voxel_filter(kf, self.cfg.voxel_leaf_submap)
            self.s2s.setInputSource(kf)
            self.s2s.calculateSourceCovariances()
            self.keyframes.append(Keyframe(pos.copy(), rot.copy(), kf, self.s2s.getSourceCovariances(), self.b.size(kf)))
        return new

    def step(self, scan_points: np.ndarray) -> Optional[FrameRecord]:
        """One LiDAR frame (registration scan in the sensor frame, Nx4 float32)."""
        t0 = time.perf_counter()
        cur = self.b.cloud(scan_points)
        if self.cfg.voxel_leaf_scan:  # preprocessPoints (odom.cc:469-474)
            cur = self.b.voxel_filter(cur, self.cfg.voxel_leaf_scan)
        n = self.b.size(cur)
        if not self._initialised:
            self._initialize_input_target(cur, n)
            return None
        # setInputSources (odom.cc:518-532)
        self.s2s.setInputSource(cur)
        self.s2m.registerInputSource(cur)
        self.b.share_source(self.s2m, self.s2s)
        # scanMatching (odom.cc:745-793)
        r1 = self.s2s.align()
        T_s2s = (self.T_s2s_prev @ r1.T.astype(np.float32)).astype(np.float32)  # propagateS2S
        self.b.hand_over_source_covs(self.s2m, self.s2s)
        self.s2s.swapSourceAndTarget()
        sel = self._submap_selection(T_s2s[:3, 3])
        changed = sel != self.submap_idx_prev
        if changed:
            submap = self.b.concat_clouds([self.keyframes[i].cloud for i in sel])
            covs = self.b.concat_covs([self.keyframes[i].covs for i in sel])
            self.s2m.setInputTarget(submap)
            self.s2m.setTargetCovariances(covs)
            self.submap_idx_prev = sel
        r2 = self.s2m.align(T_s2s)
        self.T = r2.T.astype(np.float32)
        res = self.s2m.getResiduals()
        self.T_s2s_prev = self.T.copy()  # odom.cc: T_s2s_prev_ = T_ after S2M
        new_kf = self._update_keyframes(cur, n)
        self.b.sync()
        rec = FrameRecord(T_s2s, self.T.copy(), r1.iterations, r2.iterations, r1.converged, r2.converged, new_kf, changed,
                          sum(self.keyframes[i].n for i in self.submap_idx_prev), time.perf_counter() - t0, float(np.mean(res)))
        self.records.append(rec)
        return rec


def run_sequence(backend, scans: Sequence[np.ndarray], cfg: LoopConfig = LoopConfig()) -> OdometryLoop:
    loop = OdometryLoop(backend, cfg)
    for s in scans:
        loop.step(s)
    return loop
