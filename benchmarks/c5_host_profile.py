#!/usr/bin/env python
"""Where does the host time of one C5 unit go?  Times every API call of the unit, in the main thread and in a
worker thread, with and without a device synchronisation after each call."""
import sys
import threading
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng, synth  # noqa: E402

w = synth.make_world()
scans = [synth.scan(f, 64, 1024, w) for f in range(3)]
rt = ng.Runtime(0)
S = ng.PointCloud(rt, scans[1])
T = ng.PointCloud(rt, scans[0])
rt.synchronize()
eye = np.eye(4, dtype=np.float32)
NAMES = ["clear", "xform_s", "xform_t", "setSrc(index)", "setTgt(index)", "align(+covs)"]


def one(eng, sync):
    t = []

    def lap():
        if sync:
            rt.synchronize()
        t.append(time.perf_counter())

    lap()
    eng.clearSource(); eng.clearTarget(); lap()
    s = S.transformed(eye); lap()
    g = T.transformed(eye); lap()
    eng.setInputSource(s); lap()
    eng.setInputTarget(g); lap()
    r = eng.align(); lap()
    return np.diff(t) * 1e3, r


def run(tag):
    eng = ng.NanoGICP(rt)
    for sync in (True, False):
        for _ in range(5):
            one(eng, sync)
        acc = np.zeros(len(NAMES))
        t0 = time.perf_counter()
        for _ in range(40):
            d, r = one(eng, sync)
            acc += d
        wall = (time.perf_counter() - t0) / 40 * 1e3
        print(tag, "sync " if sync else "async", " ".join(f"{n}={v / 40:.3f}" for n, v in zip(NAMES, acc)), f"sum={acc.sum() / 40:.3f} wall={wall:.3f} ms")


run("main-thread  ")
th = threading.Thread(target=run, args=("worker-thread",))
th.start()
th.join()
rt.close()
