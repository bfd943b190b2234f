"""Inputs shared by the tests that compare the segmentation stage with the reference's own DetectionModule code
(oracle/refdet.py): images larger than the window 156..356 the reference hard-codes, integer values for the ROS
parameters it reads as ints."""
import numpy as np

from dynamic_direct_lidar_odometry_b200 import synth

WINDOW = dict(window_row_min=156, window_row_max=356, window_col_min=156, window_col_max=356)

VARIANTS = [
    # the reference's defaults except for the scanner geometry
    dict(ground_rows=30, theta=60.0 / 180.0 * np.pi, valid_point_num=15, min_line_num=5, valid_line_num=5, max_delta_z=3.0, max_elevation=2.0),
    # a wide ground band, permissive segment tests: more accepted segments
    dict(ground_rows=200, theta=0.3, valid_point_num=6, min_line_num=2, valid_line_num=2, max_delta_z=6.0, max_elevation=5.0),
    # no ground removal, very permissive
    dict(ground_rows=0, theta=0.1, valid_point_num=4, min_line_num=2, valid_line_num=2, max_delta_z=10.0, max_elevation=9.0),
]


def reference_case(variant: int, frame: int = 9, size: int = 512, dropout: float = 0.02):
    """(params, world-frame organised scan, pose, residual plane)"""
    sc = synth.organized_scan(frame, size, size, dropout=dropout)
    T = synth.pose(frame).astype(np.float32)
    st = synth.organized_transform(sc, T)
    params = dict(rows=size, cols=size, ang_bottom=22, minimum_range=1, sensor_mount_angle=0, max_distance=40, **WINDOW, **VARIANTS[variant])
    res = np.abs(np.random.default_rng(frame).normal(0.0, 0.05, (size, size))).astype(np.float32)
    res[np.random.default_rng(frame + 1).random((size, size)) < 0.3] = 0.0
    return params, st, T, res


def golden_case():
    """A 360 x 360 image with returns only in a band around the window (everything else NaN / zero), so that the committed
    fixture stays small."""
    params, st, T, res = reference_case(1, frame=21, size=360, dropout=0.03)
    band = np.zeros((360, 360), dtype=bool)
    band[150:360, 150:360] = True
    st = st.copy()
    st[~band] = np.nan
    res = np.where(band, res, np.float32(0)).astype(np.float32)
    return params, st, T, res
