"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes view of oracle/_ref/libdetection_ref.so: the reference's OWN DetectionModule code on the segmentation path
(class declaration unmodified, the eight member functions extracted from src/detection/detection.cpp at build time;
see ref_detection_shim.cpp).  Used to pin oracle/oracle_segmentation.cpp and, on the GPU box, the CUDA stage itself.

Limits that are the reference's own: the `valid_range` window is hard-coded to rows/cols 156..356, and the ROS parameters
ang_bottom, groundAngleThreshold, minimumRange, sensorMountAngle and maxDistance are read with integer defaults, i.e. as
ints.  `segment` therefore refuses other windows and non-integer values for those parameters.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .pyoracle import REF_DETECTION_LIB_PATH

_lib = None
ROS_NAMES = {
    "rows": "odomNode/detection/rows", "cols": "odomNode/detection/columns", "ground_rows": "odomNode/detection/groundRows",
    "valid_point_num": "odomNode/detection/validPointNum", "min_line_num": "odomNode/detection/minLineNum",
    "valid_line_num": "odomNode/detection/validLineNum", "ang_bottom": "odomNode/detection/ang_bottom",
    "ground_angle_threshold": "odomNode/detection/groundAngleThreshold", "minimum_range": "odomNode/detection/minimumRange",
    "sensor_mount_angle": "odomNode/detection/sensorMountAngle", "theta": "odomNode/detection/theta",
    "min_delta_z": "odomNode/detection/minDeltaZ", "max_delta_z": "odomNode/detection/maxDeltaZ",
    "max_distance": "odomNode/detection/maxDistance", "max_elevation": "odomNode/detection/maxElevation",
}
INT_TYPED = ("ang_bottom", "ground_angle_threshold", "minimum_range", "sensor_mount_angle", "max_distance")
REFERENCE_WINDOW = dict(window_row_min=156, window_row_max=356, window_col_min=156, window_col_max=356)


def available() -> bool:
    return REF_DETECTION_LIB_PATH.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(REF_DETECTION_LIB_PATH))
        L.refdet_segment.restype = C.c_int
        L.refdet_segment.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.refdet_residual_cloud.restype = None
        L.refdet_residual_cloud.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def segment(params: dict, scan_t, T, residuals=None):
    """params: the keyword arguments of pyoracle.SegParams.  Same result dict as pyoracle.segment_scan (no `borderline`)."""
    for k, v in REFERENCE_WINDOW.items():
        if params.get(k, v) != v:
            raise ValueError("the reference hard-codes the window 156..356")
    if params.get("scan_in_sensor_frame", 0):
        raise ValueError("the reference's DetectionModule takes the transformed scan")
    for k in INT_TYPED:
        if k in params and float(params[k]) != int(params[k]):
            raise ValueError(f"{k} is read as an int by the reference (its ROS default is an integer literal)")
    H, W = params["rows"], params["cols"]
    items = [(ROS_NAMES[k], float(v)) for k, v in params.items() if k in ROS_NAMES]
    names = (C.c_char_p * len(items))(*[n.encode() for n, _ in items])
    values = (C.c_double * len(items))(*[v for _, v in items])
    s = np.ascontiguousarray(np.asarray(scan_t, dtype=np.float32).reshape(H * W, -1)[:, :4])
    T16 = np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).reshape(-1)
    r = None if residuals is None else np.ascontiguousarray(residuals, dtype=np.float32).reshape(-1)
    label = np.empty(H * W, dtype=np.int32)
    rng = np.empty(H * W, dtype=np.float32)
    ground = np.empty(H * W, dtype=np.int8)
    avg = np.zeros(H * W, dtype=np.float64)
    n = lib().refdet_segment(names, values, len(items), s.ctypes.data, s.ctypes.data, H * W, T16.ctypes.data, None if r is None else r.ctypes.data,
                             label.ctypes.data, rng.ctypes.data, ground.ctypes.data, avg.ctypes.data, avg.size)
    if n < 0:
        raise ValueError("image size does not match rows x cols")
    return dict(label_mat=label.reshape(H, W), range_mat=rng.reshape(H, W), ground_mat=ground.reshape(H, W), label_count=n,
                avg_residuals=avg[:n].copy())


def residual_cloud(points, residuals) -> np.ndarray:
    """the reference's own residual-cloud loop (odom.cc:804-827): (512, 512, 4) float32, angles +-60 degrees"""
    p = np.ascontiguousarray(np.asarray(points, dtype=np.float32)[:, :4])
    if p.shape[1] < 4:
        p = np.ascontiguousarray(np.concatenate([p, np.ones((len(p), 4 - p.shape[1]), dtype=np.float32)], axis=1))
    r = np.ascontiguousarray(residuals, dtype=np.float64)
    out = np.empty((512, 512, 4), dtype=np.float32)
    lib().refdet_residual_cloud(p.ctypes.data, len(p), r.ctypes.data, out.ctypes.data)
    return out
