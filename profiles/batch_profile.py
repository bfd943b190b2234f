"""A short batched run for ncu launch lists:  python profiles/batch_profile.py [units] [wave] [lanes] [c2|c5]

`ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 6000 --csv --log-file X python profiles/batch_profile.py 64`
then `python profiles/launch_summary.py X <units>` gives time and warp instructions per kernel and PER REGISTRATION
(every unit of the run is of the same kind, so the totals divide evenly)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng  # noqa: E402
from dynamic_direct_lidar_odometry_b200 import synth  # noqa: E402

units = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wave = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kind = sys.argv[4] if len(sys.argv) > 4 else "c2"
w = synth.make_world()
b = ng.Batch(0, lanes=lanes, host_threads=1, mode="waves", wave_units=wave)
if kind == "c2":
    src, tgt, guess = synth.workload_c2()
    b.set_shared_target(b.stage(tgt))
    ids = [b.stage(synth.scan(50 + f, 64, 1024, w)) for f in range(8)]
    jobs = [(ids[i % 8], -1, synth.perturbed_guess(synth.pose(50 + i % 8))) for i in range(units)]
else:
    ids = [b.stage(synth.scan(f, 64, 1024, w)) for f in range(9)]
    jobs = [(ids[i % 8 + 1], ids[i % 8], None) for i in range(units)]
t0 = time.perf_counter()
res = b.run(jobs)
dt = time.perf_counter() - t0
print(f"{kind}: {units} units, {units / dt:.0f} units/s (first run, pools cold), converged {sum(r.converged for r in res)}, "
      f"mean outer iterations {np.mean([r.iterations + 1 for r in res]):.2f}, rounds/polls {b.stats()}")
b.close()
