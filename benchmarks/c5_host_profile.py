import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng, synth
w = synth.make_world()
scans = [synth.scan(f, 64, 1024, w) for f in range(3)]
rt = ng.Runtime(0)
S = ng.PointCloud(rt, scans[1]); T = ng.PointCloud(rt, scans[0]); rt.synchronize()
eye = np.eye(4, dtype=np.float32)
eng = ng.NanoGICP(rt)
def one(sync):
    t = []
    def lap():
        if sync: rt.synchronize()
        t.append(time.perf_counter())
    lap()
    eng.clearSource(); eng.clearTarget(); lap()
    s = S.transformed(eye); lap()
    g = T.transformed(eye); lap()
    eng.setInputSource(s); lap()
    eng.setInputTarget(g); lap()
    eng.calculateSourceCovariances(); lap()
    eng.calculateTargetCovariances(); lap()
    r = eng.align(); lap()
    return np.diff(t) * 1e3, r
for sync in (True, False):
    for _ in range(5): one(sync)
    acc = np.zeros(8)
    for _ in range(20):
        d, r = one(sync); acc += d
    names = ["clear", "xform_s", "xform_t", "setSrc(index)", "setTgt(index)", "covS", "covT", "align"]
    print("sync" if sync else "async", " ".join(f"{n}={v/20:.3f}" for n, v in zip(names, acc)), f"total={acc.sum()/20:.3f} ms  iters={r.iterations} lin={r.n_linearize} err={r.n_compute_error}")
