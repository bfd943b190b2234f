"""Small cases of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python profiles/sanitize_case.py
    compute-sanitizer --tool racecheck python profiles/sanitize_case.py
    compute-sanitizer --tool synccheck python profiles/sanitize_case.py

Covers smoke() (index build incl. the 16-CTA cluster sort, k-NN, covariances, the cooperative align kernel, the
segmentation stage), the preprocessing filters, an S2S -> S2M hand-over, and a batch of units on two lanes."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as entry  # noqa: E402
from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng  # noqa: E402
from dynamic_direct_lidar_odometry_b200 import synth  # noqa: E402

entry.smoke()
w = synth.make_world()
scans = [synth.scan(f, 16, 256, w) for f in range(4)]
rt = ng.Runtime(0)
c = ng.PointCloud(rt, scans[0])
idx, d2 = c.nearestKSearch(scans[1][:200, :3], 5)
v = c.voxel_filtered(0.5)
cr = c.cropped([-1, -1, -1], [1, 1, 1], negative=True)
g = ng.NanoGICP(rt)
g.setInputSource(ng.PointCloud(rt, scans[1]))
g.setInputTarget(c)
r = g.align()
g.getResiduals()
g.residualImage(64, 64)
g.swapSourceAndTarget()
g.setInputSource(ng.PointCloud(rt, scans[2]))
r2 = g.align()
rt.set_align_blocks(5)
g.clearSource()
g.setInputSource(ng.PointCloud(rt, scans[3]))
r3 = g.align()
b = ng.Batch(0, lanes=2, host_threads=2)
ids = [b.stage(s) for s in scans]
res = b.run([(ids[i % 3 + 1], ids[i % 3], None) for i in range(6)])
b.set_shared_target(ids[0])
res2 = b.run([(ids[1], -1, None), (ids[2], -1, None)])
print("sanitize case ok:", r.converged, r2.converged, r3.converged, all(x.converged for x in res), len(v), len(cr))
b.close()
del g, c, v, cr
rt.close()
