"""Host-side mirror of the segmentation stage of the reference's DetectionModule.

Same call protocol as OdomNode::applySegmentation (odom.cc:853-857):

    det = DetectionModule(rt, rows=..., cols=..., ...)           # loadParams, detection.cpp:72-126
    det.projectScan(scan, scan_t, T, T_s2s)                      # detection.cpp:254-329
    det.projectResiduals(residuals_cloud)                        # detection.cpp:203-252
    det.applySegmentation()                                      # groundRemoval + cloudSegmentation, :191-196, :448-724
    det.label_mat, det.range_mat, det.ground_mat, det.avg_residuals, det.getSegmentsCount(), det.getGroundIndices()

The three calls only stage their arguments; the device work is one C-ABI call (ddlo_segment_scan) issued by
applySegmentation.  Bounding boxes, tracking and visualisation (computeAllObjects, trackDetections, visualize) are
not part of this stage.  There is no CPU path: without the CUDA library the first call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import binding as B

INVALID_SEGMENT = 999999  # label of rejected segments (detection.cpp:720)


class DetectionModule:
    def __init__(self, rt, **params):
        self.rt = rt
        self.params = B.SegmentationParams(**params)
        self.H_, self.W_ = self.params.rows, self.params.cols
        self._scan_t: Optional[np.ndarray] = None
        self._T = np.eye(4, dtype=np.float32)
        self._residuals: Optional[np.ndarray] = None
        self._engine = None
        self.icp_residuals_set_ = False
        self.label_mat = self.range_mat = self.ground_mat = None
        self.avg_residuals = np.zeros(0)
        self.label_count_ = 1
        self.device_ms = 0.0

    # detection.cpp:254-329: only the transformed scan and the pose enter the range image.  With cloud_in_t = None the
    # sensor-frame scan cloud_in is moved by T on the device first (OdomNode::transformScans, odom.cc:957-963).
    def projectScan(self, cloud_in, cloud_in_t, T, T_s2s=None):
        self.params.scan_in_sensor_frame = 1 if cloud_in_t is None else 0
        s = np.ascontiguousarray(cloud_in if cloud_in_t is None else cloud_in_t, dtype=np.float32)
        if s.ndim == 3:
            s = s.reshape(-1, s.shape[-1])
        if s.ndim != 2 or s.shape[0] != self.H_ * self.W_ or s.shape[1] < 3:
            raise ValueError(f"the segmentation scan must be organised: {self.H_} x {self.W_} points of >= 3 floats")
        self._scan_t = s
        self._T = np.asarray(T, dtype=np.float32).reshape(4, 4)
        self.icp_residuals_set_ = self.icp_residuals_set_ and (self._residuals is not None or self._engine is not None)

    # detection.cpp:203-252: residuals_cloud is the (rows, cols, 4) residual image (NanoGICP.residualImage), or just its intensity plane
    def projectResiduals(self, residuals_cloud):
        r = np.asarray(residuals_cloud, dtype=np.float32)
        if r.ndim == 3:
            finite = np.isfinite(r[..., :3]).all(axis=-1)
            r = np.where(finite, r[..., 3], np.float32(0))
        if r.size != self.H_ * self.W_:
            raise ValueError("the residual image must have rows x cols pixels")
        self._residuals = np.ascontiguousarray(r, dtype=np.float32).reshape(self.H_, self.W_)
        self._engine = None
        self.icp_residuals_set_ = True

    def projectResidualsFrom(self, engine, angle_min: float = -np.pi / 3, angle_max: float = np.pi / 3):
        """projectResiduals without the host round trip: the residual cloud of `engine`'s last align (odom.cc:804-827) is
        built on the device with rows x cols cells and its intensity channel used directly (ddlo_gicp_segment_scan)."""
        self._engine = (engine, float(angle_min), float(angle_max))
        self._residuals = None
        self.icp_residuals_set_ = True

    # detection.cpp:191-199 (groundRemoval + cloudSegmentation; icp_residuals_set_ is cleared at the end)
    def applySegmentation(self):
        if self._scan_t is None:
            raise RuntimeError("projectScan has not been called")
        H, W = self.H_, self.W_
        label = np.empty((H, W), dtype=np.int32)
        rng = np.empty((H, W), dtype=np.float32)
        ground = np.empty((H, W), dtype=np.int8)
        avg = np.zeros(H * W, dtype=np.float64)
        count, ms = C.c_int(0), C.c_float(0)
        T16 = np.ascontiguousarray(self._T.T)
        res = self._residuals if self.icp_residuals_set_ else None
        eng = self._engine if self.icp_residuals_set_ else None
        if eng is not None:
            B.check(B.load().ddlo_gicp_segment_scan(eng[0]._g, C.byref(self.params), B.ptr(self._scan_t), self._scan_t.shape[1] * 4, B.ptr(T16),
                                                    eng[1], eng[2], B.ptr(label), B.ptr(rng), B.ptr(ground), B.ptr(avg), avg.size,
                                                    C.byref(count), C.byref(ms)))
        else:
            B.check(B.load().ddlo_segment_scan(self.rt._h, C.byref(self.params), B.ptr(self._scan_t), self._scan_t.shape[1] * 4, B.ptr(T16),
                                               None if res is None else B.ptr(res), B.ptr(label), B.ptr(rng), B.ptr(ground), B.ptr(avg),
                                               avg.size, C.byref(count), C.byref(ms)))
        self._engine = None
        self.label_mat, self.range_mat, self.ground_mat = label, rng, ground
        self.label_count_ = count.value
        self.avg_residuals = avg[: count.value].copy()
        self.device_ms = ms.value
        self.icp_residuals_set_ = False
        return self

    # getters of the reference used by OdomNode (odom.cc:862-882, :1449)
    def getSegmentsCount(self) -> int:
        return self.label_count_ - 1

    def getGroundIndices(self) -> np.ndarray:
        return np.flatnonzero(self.ground_mat.reshape(-1) == 1).astype(np.int32)

    def getLabelIndices(self, label: int) -> np.ndarray:
        """label_indices_i_[label] (detection.cpp:528-543): raster-ordered pixel indices of one accepted segment"""
        return np.flatnonzero(self.label_mat.reshape(-1) == label).astype(np.int32)
