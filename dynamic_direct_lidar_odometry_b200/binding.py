"""ctypes view of include/ddlo_gicp.h (libddlo_gicp_b200.so).

This module only declares signatures and converts errors; it holds no algorithm.  If the shared
library is missing or no CUDA device is present every entry point raises — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

OK = 0
REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
OPT_GAUSS_NEWTON, OPT_LEVENBERG_MARQUARDT = 0, 1
FLAG_CONVERGED, FLAG_LM_FAILED, FLAG_COVS_COMPUTED = 1, 2, 4
BATCH_LANES, BATCH_WAVES = 0, 1

ERROR_NAMES = {-1: "INVALID", -2: "CUDA", -3: "EMPTY", -4: "TOO_FEW", -5: "NOT_READY", -6: "SIZE", -7: "NONFINITE", -8: "UNSUPPORTED"}


class DdloError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"ddlo_gicp: {ERROR_NAMES.get(code, code)}: {text}")
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("k_correspondences", C.c_int),
        ("regularization_method", C.c_int),
        ("max_iterations", C.c_int),
        ("optimizer", C.c_int),
        ("lm_max_iterations", C.c_int),
        ("reserved_", C.c_int),
        ("max_correspondence_distance", C.c_double),
        ("transformation_epsilon", C.c_double),
        ("rotation_epsilon", C.c_double),
        ("lm_init_lambda_factor", C.c_double),
    ]


class AlignResult(C.Structure):
    _fields_ = [
        ("final_transformation", C.c_float * 16),
        ("final_hessian", C.c_double * 36),
        ("flags", C.c_int),
        ("nr_iterations", C.c_int),
        ("n_linearize", C.c_int),
        ("n_compute_error", C.c_int),
        ("final_error", C.c_double),
        ("lm_lambda", C.c_double),
    ]


_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)
_ip = C.POINTER(C.c_int)

# name -> argtypes; every function returns int unless listed in _SPECIAL
SIGNATURES = {
    "ddlo_device_count": [_ip],
    "ddlo_runtime_create": [C.c_int, _vpp],
    "ddlo_runtime_destroy": [_vp],
    "ddlo_runtime_synchronize": [_vp],
    "ddlo_runtime_set_align_blocks": [_vp, C.c_int],
    "ddlo_runtime_timer_begin": [_vp],
    "ddlo_runtime_timer_end": [_vp, C.POINTER(C.c_float)],
    "ddlo_runtime_launch_count": [_vp, C.POINTER(C.c_longlong)],
    "ddlo_runtime_flush_l2": [_vp, C.c_size_t],
    "ddlo_runtime_event_record": [_vp, C.c_int],
    "ddlo_runtime_event_elapsed": [_vp, C.c_int, C.c_int, C.POINTER(C.c_float)],
    "ddlo_host_alloc": [C.c_size_t, _vpp],
    "ddlo_host_free": [_vp],
    "ddlo_cloud_create": [_vp, _vp, C.c_int, C.c_int, _vpp],
    "ddlo_cloud_create_from_device": [_vp, _vp, C.c_int, _vpp],
    "ddlo_cloud_retain": [_vp],
    "ddlo_cloud_release": [_vp],
    "ddlo_cloud_size": [_vp, _ip],
    "ddlo_cloud_download": [_vp, _vp],
    "ddlo_cloud_build_index": [_vp],
    "ddlo_cloud_has_index": [_vp, _ip],
    "ddlo_cloud_knn": [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp],
    "ddlo_cloud_transform": [_vp, _vp, _vpp],
    "ddlo_cloud_concat": [_vp, _vpp, C.c_int, _vpp],
    "ddlo_cloud_voxel_filter": [_vp, C.c_float, C.c_float, C.c_float, _vpp],
    "ddlo_cloud_crop_box": [_vp, _vp, _vp, C.c_int, C.c_int, _vpp],
    "ddlo_cloud_extract_stride": [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vpp],
    "ddlo_covs_compute": [_vp, C.c_int, C.c_int, _vpp],
    "ddlo_covs_from_host": [_vp, _vp, C.c_int, _vpp],
    "ddlo_covs_to_host": [_vp, _vp],
    "ddlo_covs_size": [_vp, _ip],
    "ddlo_covs_retain": [_vp],
    "ddlo_covs_release": [_vp],
    "ddlo_covs_concat": [_vp, _vpp, C.c_int, _vpp],
    "ddlo_gicp_create": [_vp, _vpp],
    "ddlo_gicp_destroy": [_vp],
    "ddlo_params_default": [C.POINTER(Params)],
    "ddlo_gicp_set_params": [_vp, C.POINTER(Params)],
    "ddlo_gicp_get_params": [_vp, C.POINTER(Params)],
    "ddlo_gicp_set_input_source": [_vp, _vp, C.c_int],
    "ddlo_gicp_set_input_target": [_vp, _vp],
    "ddlo_gicp_clear_source": [_vp],
    "ddlo_gicp_clear_target": [_vp],
    "ddlo_gicp_set_source_covariances": [_vp, _vp],
    "ddlo_gicp_set_target_covariances": [_vp, _vp],
    "ddlo_gicp_get_source_covariances": [_vp, _vpp],
    "ddlo_gicp_get_target_covariances": [_vp, _vpp],
    "ddlo_gicp_get_input_source": [_vp, _vpp],
    "ddlo_gicp_get_input_target": [_vp, _vpp],
    "ddlo_gicp_calculate_source_covariances": [_vp],
    "ddlo_gicp_calculate_target_covariances": [_vp],
    "ddlo_gicp_swap_source_and_target": [_vp],
    "ddlo_gicp_align": [_vp, _vp, C.POINTER(AlignResult)],
    "ddlo_gicp_align_async": [_vp, _vp],
    "ddlo_gicp_align_finish": [_vp, C.POINTER(AlignResult)],
    "ddlo_gicp_aligned_cloud": [_vp, _vpp],
    "ddlo_gicp_linearize": [_vp, _vp, _vp, _vp, C.POINTER(C.c_double)],
    "ddlo_gicp_compute_error": [_vp, _vp, C.POINTER(C.c_double)],
    "ddlo_gicp_get_correspondences": [_vp, _vp, _vp, C.c_int],
    "ddlo_gicp_get_mahalanobis": [_vp, _vp, C.c_int],
    "ddlo_gicp_get_residuals": [_vp, _vp, C.c_int],
    "ddlo_gicp_get_residuals_async": [_vp, _vp, C.c_int],
    "ddlo_gicp_get_residual_vectors": [_vp, _vp, _vp, C.c_int],
    "ddlo_gicp_residual_image": [_vp, C.c_int, C.c_int, C.c_double, C.c_double, _vp],
    "ddlo_segment_scan": [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _ip, _vp],
    "ddlo_gicp_segment_scan": [_vp, _vp, _vp, C.c_int, _vp, C.c_double, C.c_double, _vp, _vp, _vp, _vp, C.c_int, _ip, _vp],
    "ddlo_gicp_align_batch": [_vpp, C.c_int, _vp, C.POINTER(AlignResult)],
    "ddlo_keyframes_create": [_vp, _vpp],
    "ddlo_keyframes_destroy": [_vp],
    "ddlo_keyframes_count": [_vp, _ip],
    "ddlo_keyframes_add": [_vp, _vp, _vp, _vp, _vp],
    "ddlo_keyframes_get": [_vp, C.c_int, _vp, _vp, _vpp, _vpp],
    "ddlo_keyframes_is_new": [_vp, _vp, _vp, C.c_float, C.c_float, _ip, _ip, C.POINTER(C.c_float), C.POINTER(C.c_float)],
    "ddlo_keyframes_get_submap": [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_double, _ip, _vpp, _vpp, _vp, C.c_int, _ip],
    "ddlo_keyframes_hulls": [_vp, _vp, _ip, _vp, _ip, C.c_int, _ip],
    "ddlo_cloud_share": [_vp],
    "ddlo_covs_share": [_vp, _vp],
    "ddlo_batch_create": [C.c_int, C.c_int, C.c_int, C.c_int, _vpp],
    "ddlo_batch_destroy": [_vp],
    "ddlo_batch_info": [_vp, _ip, _ip, _ip],
    "ddlo_batch_set_params": [_vp, C.POINTER(Params)],
    "ddlo_batch_set_mode": [_vp, C.c_int, C.c_int],
    "ddlo_batch_stats": [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
    "ddlo_batch_stage_cloud": [_vp, _vp, C.c_int, C.c_int, _ip],
    "ddlo_batch_staged_count": [_vp, _ip],
    "ddlo_batch_set_shared_target": [_vp, C.c_int, _vp],
    "ddlo_batch_submit": [_vp, _vp, C.c_int, _vp],
    "ddlo_batch_submit_host": [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp],
    "ddlo_batch_wait": [_vp],
    "ddlo_batch_run": [_vp, _vp, C.c_int, _vp],
    "ddlo_batch_launch_count": [_vp, C.POINTER(C.c_longlong)],
}


class BatchJob(C.Structure):
    """struct ddlo_batch_job (include/ddlo_gicp.h)"""

    _fields_ = [("source", C.c_int), ("target", C.c_int), ("guess", C.c_float * 16)]

class SegmentationParams(C.Structure):
    """struct ddlo_segmentation_params (include/ddlo_gicp.h); the defaults are the reference's (detection.cpp:76-105, :520-522)."""

    _fields_ = [(n, C.c_int) for n in ("rows", "cols", "ground_rows", "valid_point_num", "min_line_num", "valid_line_num",
                                       "window_row_min", "window_row_max", "window_col_min", "window_col_max", "scan_in_sensor_frame",
                                       "unordered_residual_sums")] + \
               [(n, C.c_float) for n in ("ang_bottom", "ground_angle_threshold", "minimum_range", "sensor_mount_angle", "theta",
                                         "min_delta_z", "max_delta_z", "max_distance", "max_elevation")]

    DEFAULTS = dict(rows=128, cols=1024, ground_rows=30, valid_point_num=15, min_line_num=5, valid_line_num=5,
                    window_row_min=156, window_row_max=356, window_col_min=156, window_col_max=356, scan_in_sensor_frame=0, unordered_residual_sums=0,
                    ang_bottom=45.0, ground_angle_threshold=10.0, minimum_range=10.0, sensor_mount_angle=10.0,
                    theta=60.0 / 180.0 * 3.14159265358979323846, min_delta_z=0.1, max_delta_z=3.0, max_distance=20.0, max_elevation=2.0)

    def __init__(self, **kw):
        super().__init__()
        unknown = set(kw) - set(self.DEFAULTS)
        if unknown:
            raise TypeError(f"unknown segmentation parameters: {sorted(unknown)}")
        for k, v in {**self.DEFAULTS, **kw}.items():
            setattr(self, k, v)


_SPECIAL = {
    "ddlo_abi_version": (C.c_int, []),
    "ddlo_last_error": (C.c_char_p, []),
}
# host-callable copies of the device math, declared in include/ddlo_gicp_testing.h
TESTING_SIGNATURES = {
    "ddlo_math_sym3_eig": [_vp, _vp, _vp],
    "ddlo_math_regularize": [_vp, C.c_int, _vp],
    "ddlo_math_ldlt6_solve": [_vp, _vp, _vp],
    "ddlo_math_ldlt6_solve_fast": [_vp, _vp, _vp],
    "ddlo_math_so3_exp": [_vp, _vp],
    "ddlo_math_sym3_inverse": [_vp, _vp],
}

_lib = None


def library_path() -> Path:
    """The in-tree library; DDLO_GICP_LIB selects another build of the same sources (kernel tuning experiments)."""
    import os

    override = os.environ.get("DDLO_GICP_LIB")
    return Path(override) if override else _build.LIB_PATH


def load() -> C.CDLL:
    """dlopen the CUDA library.  Raises if it has not been built — the product has no other path."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); there is no CPU fallback")
    L = C.CDLL(str(path))
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    for name, (res, args) in _SPECIAL.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    for name, args in TESTING_SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = None
        fn.argtypes = args
    L.ddlo_gicp_debug_block_times.restype = C.c_int
    L.ddlo_gicp_debug_block_times.argtypes = [_vp, _vp, C.c_int]
    L.ddlo_gicp_debug_visits.restype = C.c_int
    L.ddlo_gicp_debug_visits.argtypes = [_vp, _vp, C.c_int]
    L.ddlo_gicp_debug_timeline.restype = C.c_int
    L.ddlo_gicp_debug_timeline.argtypes = [_vp, _vp, C.c_int]
    for name, args in (("ddlo_hull_convex", [_vp, C.c_int, _vp, C.c_int]), ("ddlo_hull_concave", [_vp, C.c_int, C.c_double, _vp, C.c_int]),
                       ("ddlo_hull_dimension", [_vp, C.c_int])):
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    L.ddlo_align_d2h_bytes.restype = C.c_int
    L.ddlo_align_d2h_bytes.argtypes = []
    L.ddlo_gicp_debug_enable.restype = C.c_int
    L.ddlo_gicp_debug_enable.argtypes = [_vp, C.c_int]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != OK:
        raise DdloError(rc, load().ddlo_last_error().decode(errors="replace"))


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)
