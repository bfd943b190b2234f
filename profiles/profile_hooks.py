"""C2 workload, stepwise hooks only (linearize / compute_error kernels in isolation) for ncu launch lists."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng  # noqa: E402
from dynamic_direct_lidar_odometry_b200 import synth  # noqa: E402

src, tgt, guess = synth.workload_c2()
rt = ng.Runtime(0)
eng = ng.NanoGICP(rt)
target = ng.PointCloud(rt, tgt)
eng.setInputTarget(target)
eng.calculateTargetCovariances()
eng.setInputSource(ng.PointCloud(rt, src))
eng.calculateSourceCovariances()
T = guess.astype(np.float64)
for _ in range(4):
    e, H, b = eng.linearize(T)
    e2 = eng.compute_error(T)
r = eng.align(guess)
print("err", e, e2, "iterations", r.iterations)
del eng, target
rt.close()
