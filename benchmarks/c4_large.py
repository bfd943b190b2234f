#!/usr/bin/env python
"""BASELINE config C4: S2M registration of a 128x2048 scan (~262k points) against a 2M-point submap.

    python benchmarks/c4_large.py [--steps 20]

Same step as bench.py (source index + source covariances k=20 + LM align, target resident), device-timed
with CUDA events after an L2 flush, plus the one-off target preparation (2M-point index + covariances).
(Parity at this size is a test: tests/test_gpu_parity.py::test_c4_properties / test_c4_align_matches_oracle.)
"""
from __future__ import annotations

import argparse
import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng
    from dynamic_direct_lidar_odometry_b200 import synth

    src, tgt, guess = synth.workload_c4()
    rt = ng.Runtime(0)
    eng = ng.NanoGICP(rt)
    target = ng.PointCloud(rt, tgt)
    # the same work once untimed (kernel modules, allocator pool), then on a fresh handle with the clock on
    warm = target.transformed(np.eye(4, dtype=np.float32))
    warm.build_index()
    ng.Covariances.compute(warm, 20)
    del warm
    rt.synchronize()
    rt.event_record(0)
    eng.setInputTarget(target)
    rt.event_record(1)
    eng.calculateTargetCovariances()
    rt.event_record(2)
    rt.synchronize()
    tgt_index_ms, tgt_cov_ms = rt.event_elapsed(0, 1), rt.event_elapsed(1, 2)
    resident = ng.PointCloud(rt, src)
    eye = np.eye(4, dtype=np.float32)

    def step():
        fresh = resident.transformed(eye)
        rt.flush_l2(256 << 20)
        rt.event_record(0)
        eng.setInputSource(fresh)
        rt.event_record(1)
        eng.calculateSourceCovariances()
        rt.event_record(2)
        eng.align_async(guess)
        rt.event_record(3)
        info = eng.align_finish()
        t = [rt.event_elapsed(0, 1), rt.event_elapsed(1, 2), rt.event_elapsed(2, 3), rt.event_elapsed(0, 3)]
        eng.clearSource()
        return t, info

    for _ in range(3):
        step()
    ts = []
    for _ in range(args.steps):
        t, info = step()
        ts.append(t)
    m = [statistics.mean(x[i] for x in ts) for i in range(4)]
    L, E, ns = info.n_linearize, info.n_compute_error, len(src)
    line = {"metric": "c4_s2m_ms_per_scan", "unit": "ms", "value": m[3], "registrations_per_s": 1e3 / m[3], "steps": args.steps,
            "source_points": ns, "target_points": len(tgt),
            "stages_ms": {"source_index": m[0], "source_covariances": m[1], "align": m[2]},
            "target_once_ms": {"index": tgt_index_ms, "covariances": tgt_cov_ms},
            "align": {"converged": info.converged, "outer_iterations": info.iterations + 1, "n_linearize": L, "n_compute_error": E},
            "algorithmic_bytes_align": (184 * L + 84 * E) * ns,
            "translation_error_vs_truth_m": float(np.abs(info.T[:3, 3] - synth.pose(50)[:3, 3]).max()),
            "l2": "flushed before every timed step (256 MiB write)"}
    print(json.dumps(line))
    del eng, target, resident
    rt.close()


if __name__ == "__main__":
    main()
