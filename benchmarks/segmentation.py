"""Range-image segmentation stage on synthetic organised scans.

For each image size: the device time of the kernels alone (CUDA events inside ddlo_segment_scan) and the time of the
whole C-ABI call with host buffers on both sides.  The CPU yardstick for the same inputs (the oracle restatement, one
core: the reference's flood fill is sequential) is timed by tests/time_segmentation_oracle.py -- only tests/ may run
the oracle.

    python benchmarks/segmentation.py [--repeats 50] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng  # noqa: E402
from dynamic_direct_lidar_odometry_b200 import synth  # noqa: E402
from dynamic_direct_lidar_odometry_b200.detection import DetectionModule  # noqa: E402

# algorithmic bytes per pixel: point 16 B + residual 4 B in, range 4 B + ground 1 B + label 4 B out
BYTES_PER_PIXEL = 29


def case(frame, beams, cols, **over):
    sc = synth.organized_scan(frame, beams, cols, dropout=0.02)
    T = synth.pose(frame).astype(np.float32)
    st = synth.organized_transform(sc, T)
    params = dict(rows=beams, cols=cols, ground_rows=beams * 3 // 8, window_row_min=0, window_row_max=beams - 1, window_col_min=0,
                  window_col_max=cols - 1, ang_bottom=22.5, minimum_range=1.0, sensor_mount_angle=0.0, max_distance=40.0)
    params.update(over)
    res = np.abs(np.random.default_rng(frame).normal(0.0, 0.05, (beams, cols))).astype(np.float32)
    return params, st, T, res


CASES = {
    "64x1024 full image": (3, 64, 1024, {}),
    "128x2048 full image": (12, 128, 2048, {}),
    "512x512, reference window 156..356": (9, 512, 512, dict(window_row_min=156, window_row_max=356, window_col_min=156, window_col_max=356,
                                                               ground_rows=30)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeats", type=int, default=50)
    ap.add_argument("--json", type=str, default=None)
    ap.add_argument("--unordered", action="store_true", help="opt-in unordered_residual_sums mode (averages to float rounding, no queue replay)")
    a = ap.parse_args()
    rt = ng.Runtime(0)
    rows = []
    for name, (frame, beams, cols, over) in CASES.items():
        params, st, T, res = case(frame, beams, cols, **over)
        if a.unordered:
            params["unordered_residual_sums"] = 1
        det = DetectionModule(rt, **params)
        dev, e2e = [], []
        for i in range(a.repeats + 5):
            t0 = time.perf_counter()
            det.projectScan(None, st, T)
            det.projectResiduals(res)
            det.applySegmentation()
            t1 = time.perf_counter()
            if i >= 5:
                dev.append(det.device_ms)
                e2e.append((t1 - t0) * 1e3)
        row = {"case": name, "unordered_residual_sums": bool(a.unordered), "pixels": beams * cols, "segments": det.getSegmentsCount(),
               "largest_segment_px": int(np.bincount(det.label_mat[(det.label_mat > 0) & (det.label_mat < 999999)].ravel()).max()) if det.getSegmentsCount() else 0,
               "rejected_px": int((det.label_mat == 999999).sum()),
               "device_ms_median": statistics.median(dev), "device_ms_min": min(dev), "e2e_ms_median": statistics.median(e2e),
               "algorithmic_GBps": beams * cols * BYTES_PER_PIXEL / (statistics.median(dev) * 1e-3) / 1e9}
        rows.append(row)
        print(json.dumps(row))
    if a.json:
        Path(a.json).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
