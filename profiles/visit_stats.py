"""Search statistics of the align kernel on the C2 step (needs a -DDDLO_VISIT_STATS build):
    DDLO_NVCC_EXTRA=-DDDLO_VISIT_STATS python -m dynamic_direct_lidar_odometry_b200.build && python profiles/visit_stats.py
"""
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
r = eng.align(guess)
v = eng.debug_visits()
if v is None:
    print("library built without -DDDLO_VISIT_STATS")
    sys.exit(0)
pct = [0, 25, 50, 75, 90, 99, 100]
for p in range(min(4, r.n_linearize)):
    nv, nl, ws = v[p, :, 0], v[p, :, 1], v[p, :, 2]
    print(f"pass {p}: node visits/query mean {nv.mean():.2f} pct{pct} {np.percentile(nv, pct)}")
    print(f"         leaf scans/query  mean {nl.mean():.2f} pct {np.percentile(nl, pct)}")
    print(f"         lock-step warp steps (per warp of 16 queries) mean {ws.mean():.2f} pct {np.percentile(ws, pct)}")
    print(f"         lane efficiency = mean visits / mean warp steps = {nv.mean() / max(ws.mean(), 1e-9):.2f}")
del eng, target
rt.close()
