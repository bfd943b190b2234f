"""C2 step with the covariances computed inside align() (nano_gicp_impl.hpp:186-193): setInputSource + align.
DDLO_PASS0_OVERLAP=1 turns the early correspondence search on (A/B of the overlap; off by default)."""
import os
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
resident = ng.PointCloud(rt, src)
eye = np.eye(4, dtype=np.float32)
ts = []
for it in range(120):
    fresh = resident.transformed(eye)
    rt.flush_l2(256 << 20)
    rt.event_record(0)
    eng.setInputSource(fresh)
    rt.event_record(1)
    eng.align_async(guess)
    rt.event_record(2)
    info = eng.align_finish()
    if it >= 20:
        ts.append((rt.event_elapsed(0, 1), rt.event_elapsed(1, 2), rt.event_elapsed(0, 2)))
    eng.clearSource()
t = np.array(ts)
print(f"overlap {'on' if os.environ.get('DDLO_PASS0_OVERLAP') else 'off'}: index {t[:, 0].mean():.4f} ms, covariances + align {t[:, 1].mean():.4f} ms, "
      f"step {t[:, 2].mean():.4f} ms (p50 {np.percentile(t[:, 2], 50):.4f}, p99 {np.percentile(t[:, 2], 99):.4f}); iterations {info.iterations + 1}, covs computed {info.covs_computed}")
print("T", info.T.ravel().tolist())
del eng, target, resident
rt.close()
