#!/usr/bin/env python
"""BASELINE config C3: synthetic 100-frame 64-beam sequence, full S2S -> S2M odometry loop on the device.

    python benchmarks/c3_sequence.py [--frames 100] [--beams 64] [--cols 1024]

Per frame (odometry_loop.py, the call sequence of OdomNode, odom.cc:518-532, 745-793, 1067-1150, 1215-1315):
upload the scan, build its index, S2S align against the previous scan (covariances of both reused
through swapSourceAndTarget), hand the source covariances to S2M, S2M align against the keyframe
submap with the S2S pose as guess, read the residuals back, and, about every metre, add a keyframe
(transform + index + covariances) and rebuild the submap by device-side concatenation.
Reports host wall-clock ms per frame (mean / p50 / p99 / max) and the drift against the generator's ground
truth as one JSON line.  (Parity of this loop with the CPU oracle is a test: tests/test_gpu_parity.py::
test_sequence_loop_matches_oracle.)
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main_cpp(args):
    import subprocess
    import tempfile

    from dynamic_direct_lidar_odometry_b200 import synth

    exe = ROOT / "tests" / "cpp" / "_build" / "odometry_sequence"
    if not exe.exists():
        sys.path.insert(0, str(ROOT))
        import __graft_entry__ as ge

        ge.build_cpp_tests()
    w = synth.make_world()
    scans = [synth.scan(f, args.beams, args.cols, w) for f in range(args.frames)]
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "scans.bin"
        with open(path, "wb") as fh:
            fh.write(np.int32(len(scans)).tobytes())
            for s in scans:
                fh.write(np.int32(len(s)).tobytes())
                fh.write(np.ascontiguousarray(s, dtype=np.float32).tobytes())
        out = subprocess.run([str(exe), str(path), str(args.k), "1.0", "15", "10", str(args.kcv), str(args.kcc), str(args.voxel), str(args.voxel), "8"],
                             capture_output=True, text=True, timeout=1200)
    if out.returncode != 0:
        raise RuntimeError(out.stderr)
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith("frame")]
    ms = np.array([float(r[9]) for r in rows])
    inv0 = np.linalg.inv(synth.pose(0))
    err_t = [float(np.abs(np.array(r[11:27], dtype=np.float64).reshape(4, 4)[:3, 3] - (inv0 @ synth.pose(int(r[1])))[:3, 3]).max()) for r in rows]
    summary = [l.split() for l in out.stdout.splitlines() if l.startswith("summary")][0]
    slow = np.argsort(-ms)[:6]
    print(json.dumps({
        "metric": "c3_odometry_loop_ms_per_frame", "unit": "ms", "frames": args.frames, "scan": f"{args.beams}x{args.cols}",
        "driver": "C++ (tests/cpp/odometry_sequence.cpp: C ABI + ddlo_keyframes_*, no Python in the frame)",
        "mean_ms": float(ms.mean()), "p50_ms": float(np.percentile(ms, 50)), "p99_ms": float(np.percentile(ms, 99)), "max_ms": float(ms.max()),
        "frames_per_s": float(1e3 / ms.mean()), "keyframes": int(summary[4]), "submap_points_last": int(rows[-1][8]),
        "submap_rebuilds": int(sum(int(r[7]) for r in rows)), "all_converged": bool(all(int(r[4]) and int(r[5]) for r in rows)),
        "s2s_iterations_mean": float(np.mean([int(r[2]) + 1 for r in rows])), "s2m_iterations_mean": float(np.mean([int(r[3]) + 1 for r in rows])),
        "final_translation_error_vs_truth_m": err_t[-1], "max_translation_error_vs_truth_m": max(err_t),
        "k_correspondences": args.k, "voxel_leaf_m": args.voxel, "submap_selection": {"knn": 10, "kcv": args.kcv, "kcc": args.kcc},
        "timer": "host wall clock per frame inside the C++ program, scan upload and residual read-back included",
        "slowest_frames": [{"frame": int(rows[i][1]), "ms": float(ms[i]), "new_keyframe": bool(int(rows[i][6])), "submap_changed": bool(int(rows[i][7])),
                            "submap_points": int(rows[i][8])} for i in slow]}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--beams", type=int, default=64)
    ap.add_argument("--cols", type=int, default=1024)
    ap.add_argument("--voxel", type=float, default=0.0, help="leaf size of the scan and keyframe voxel filters on the device (0 = off; the DLO yaml uses 0.25 - 0.5 m)")
    ap.add_argument("--segmentation", action="store_true",
                    help="also run the range-image segmentation stage (OdomNode::applySegmentation, odom.cc:853-857) on every frame")
    ap.add_argument("--k", type=int, default=20, help="kCorrespondences of both engines (engine default 20; the DLO yaml uses 10)")
    ap.add_argument("--cpp", action="store_true", help="run the frame loop in C++ (tests/cpp/odometry_sequence.cpp on the C ABI and the "
                    "device-resident keyframe store): no Python inside a frame")
    ap.add_argument("--kcv", type=int, default=10, help="--cpp: nearest convex-hull keyframes added to the submap (the reference's default)")
    ap.add_argument("--kcc", type=int, default=10, help="--cpp: nearest concave-hull keyframes")
    args = ap.parse_args()
    if args.cpp:
        return main_cpp(args)

    from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng
    from dynamic_direct_lidar_odometry_b200 import odometry_loop as ol
    from dynamic_direct_lidar_odometry_b200 import synth

    w = synth.make_world()
    t0 = time.perf_counter()
    organized = None
    if args.segmentation:  # the registration scan is the organised scan without its empty pixels
        organized = [synth.organized_scan(f, args.beams, args.cols, w) for f in range(args.frames)]
        scans = [np.ascontiguousarray(o.reshape(-1, 4)[np.isfinite(o[..., 0]).reshape(-1)]) for o in organized]
    else:
        scans = [synth.scan(f, args.beams, args.cols, w) for f in range(args.frames)]
    gen_s = time.perf_counter() - t0
    cfg = ol.LoopConfig(k_correspondences_s2s=args.k, k_correspondences_s2m=args.k,
                        voxel_leaf_scan=args.voxel or None, voxel_leaf_submap=args.voxel or None)

    rt = ng.Runtime(0)
    ol.run_sequence(ol.GpuBackend(rt), scans[: min(8, args.frames)], cfg)  # warm-up: allocator pools, first launches
    rt.synchronize()
    seg_ms, seg_dev_ms, seg_counts = [], [], []
    if not args.segmentation:
        loop = ol.run_sequence(ol.GpuBackend(rt), scans, cfg)
    else:
        from dynamic_direct_lidar_odometry_b200.detection import DetectionModule

        det = DetectionModule(rt, rows=args.beams, cols=args.cols, ground_rows=args.beams * 3 // 8, window_row_min=0, window_row_max=args.beams - 1,
                              window_col_min=0, window_col_max=args.cols - 1, ang_bottom=22.5, minimum_range=1.0, sensor_mount_angle=0.0,
                              max_distance=40.0)
        loop = ol.OdometryLoop(ol.GpuBackend(rt), cfg)
        for f, scan in enumerate(scans):
            rec = loop.step(scan)
            if rec is None:
                continue
            # segmentation_scan_t_ = the organised scan moved by the frame's pose (odom.cc:957-963; host side, not timed)
            scan_t = synth.organized_transform(organized[f], rec.T)
            residuals = loop.s2m.getResiduals()
            t1 = time.perf_counter()
            plane = np.zeros(args.beams * args.cols, dtype=np.float32)  # residual per pixel of the organised scan (odom.cc:804-827)
            plane[np.isfinite(organized[f][..., 0]).reshape(-1)] = residuals
            det.projectScan(None, scan_t, rec.T)
            det.projectResiduals(plane)
            det.applySegmentation()
            seg_ms.append((time.perf_counter() - t1) * 1e3)
            seg_dev_ms.append(det.device_ms)
            seg_counts.append(det.getSegmentsCount())
    ms = np.array([r.seconds for r in loop.records]) * 1e3
    inv0 = np.linalg.inv(synth.pose(0))
    err_t = [float(np.abs(r.T[:3, 3].astype(np.float64) - (inv0 @ synth.pose(f))[:3, 3]).max()) for f, r in enumerate(loop.records, start=1)]
    line = {
        "metric": "c3_odometry_loop_ms_per_frame", "unit": "ms", "frames": args.frames, "scan": f"{args.beams}x{args.cols}",
        "mean_ms": float(ms.mean()), "p50_ms": float(np.percentile(ms, 50)), "p99_ms": float(np.percentile(ms, 99)), "max_ms": float(ms.max()),
        "frames_per_s": float(1e3 / ms.mean()), "keyframes": len(loop.keyframes),
        "submap_points_last": loop.records[-1].submap_points, "submap_rebuilds": int(sum(r.submap_changed for r in loop.records)),
        "all_converged": bool(all(r.s2s_converged and r.s2m_converged for r in loop.records)),
        "s2s_iterations_mean": float(np.mean([r.s2s_iterations + 1 for r in loop.records])),
        "s2m_iterations_mean": float(np.mean([r.s2m_iterations + 1 for r in loop.records])),
        "final_translation_error_vs_truth_m": err_t[-1], "max_translation_error_vs_truth_m": max(err_t),
        "k_correspondences": args.k, "voxel_leaf_m": args.voxel, "timer": "host wall clock per frame, scan upload and residual read-back included",
        "scan_generation_s": gen_s,
        **({"segmentation_ms_per_frame_mean": float(np.mean(seg_ms)), "segmentation_ms_per_frame_p99": float(np.percentile(seg_ms, 99)),
            "segmentation_device_ms_mean": float(np.mean(seg_dev_ms)), "segments_per_frame_mean": float(np.mean(seg_counts)),
            "segmentation_timer": "host wall clock around projectScan + projectResiduals + applySegmentation (host buffers both sides)"}
           if seg_ms else {}),
        "slowest_frames": [{"frame": int(i) + 1, "ms": float(ms[i]), "new_keyframe": bool(loop.records[i].new_keyframe),
                            "submap_changed": bool(loop.records[i].submap_changed), "submap_points": int(loop.records[i].submap_points)}
                           for i in np.argsort(-ms)[:6]],
    }
    print(json.dumps(line))
    del loop
    rt.close()


if __name__ == "__main__":
    main()
