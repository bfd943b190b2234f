"""One C2 step (index + covariances + align) for ncu captures:  python profiles/profile_step.py [reps]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng  # noqa: E402
from dynamic_direct_lidar_odometry_b200 import synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
src, tgt, guess = synth.workload_c2()
rt = ng.Runtime(0)
eng = ng.NanoGICP(rt)
eng.debug_enable(True)  # the kernel records its phase timeline only when asked to
target = ng.PointCloud(rt, tgt)
eng.setInputTarget(target)
eng.calculateTargetCovariances()
for _ in range(reps):
    eng.clearSource()
    eng.setInputSource(ng.PointCloud(rt, src))
    eng.calculateSourceCovariances()
    r = eng.align(guess)
prev = 0.0
for tag, us in eng.debug_timeline():
    print(f"  {tag:12s} {us:9.1f} us  (+{us - prev:7.1f})")
    prev = us
bt = eng.debug_block_times()
for p in range(min(r.n_linearize, 8)):
    st, sd, bd, sy = bt[p, :, 0], bt[p, :, 1], bt[p, :, 2], bt[p, :, 3]
    print(f"  pass {p}: start {st.min():7.1f}..{st.max():7.1f}  search {np.percentile(sd - st, [0, 50, 100]).round(1)}  "
          f"phaseB {np.percentile(bd - sd, [0, 50, 100]).round(1)}  wait {np.percentile(sy - bd, [0, 50, 100]).round(1)}  "
          f"sync2-search {np.percentile(bt[p, :, 6] - sd, [0, 50, 100]).round(1)} final {np.percentile(bt[p, :, 7] - bt[p, :, 6], [0, 50, 100]).round(1)} "
          f"lin_point max/warp {np.percentile(bt[p, :, 4], [0, 50, 100]).round(1)} min/warp {np.percentile(bt[p, :, 5], [0, 50, 100]).round(1)}")
print("iterations", r.iterations, "converged", r.converged, "lin", r.n_linearize, "err", r.n_compute_error)
del eng, target
rt.close()
