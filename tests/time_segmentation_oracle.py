"""CPU yardstick for benchmarks/segmentation.py: the oracle restatement of the segmentation stage timed on the same
synthetic inputs (one core; the reference's flood fill is sequential).  A script, not a test:

    python tests/time_segmentation_oracle.py
"""
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "benchmarks"))

from oracle import pyoracle as po  # noqa: E402

import segmentation as seg_bench  # noqa: E402  (benchmarks/segmentation.py: the shared case generator)

if __name__ == "__main__":
    po.build()
    for name, (frame, beams, cols, over) in seg_bench.CASES.items():
        params, st, T, res = seg_bench.case(frame, beams, cols, **over)
        sp = po.SegParams(**params)
        t = []
        for _ in range(7):
            t0 = time.perf_counter()
            o = po.segment_scan(sp, st, T, res)
            t.append((time.perf_counter() - t0) * 1e3)
        print(json.dumps({"case": name, "cpu_oracle_ms_median": statistics.median(t), "cpu_oracle_ms_min": min(t), "cores": 1,
                          "segments": o["label_count"] - 1}))
