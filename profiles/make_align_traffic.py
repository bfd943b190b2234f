"""profiles/align_traffic.json from an `ncu --set full` report that holds a k_align launch:
    python profiles/make_align_traffic.py gpurun_out/r02_prof_step.ncu-rep
The file is stamped with a fingerprint of the kernel's sources (bench.kernel_source_sha16); bench.py quotes the
traffic only while the fingerprint matches the sources it runs with."""
import csv
import datetime
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
pick = [r for r in rows[2:] if "k_align" in r[idx["Kernel Name"]]]
if not pick:
    raise SystemExit("no k_align launch in the report")
r = pick[-1]


def val(name):
    return float(r[idx[name]].replace(",", ""))


def to_bytes(name):
    unit = rows[1][idx[name]].strip().lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
    return val(name) * scale


rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
dur_unit = rows[1][idx["gpu__time_duration.sum"]].strip().lower()
dur = val("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(dur_unit, 1.0)
doc = {
    "kernel": "k_align",
    "dram_bytes_per_launch": rd + wr,
    "dram_bytes_read": rd,
    "dram_bytes_write": wr,
    "duration_us_under_ncu": dur,
    "kernel_source_sha16": bench.kernel_source_sha16(),
    "captured": datetime.date.today().isoformat() + f", ncu --set full --clock-control none, C2 workload, one launch of profiles/profile_step.py ({Path(rep).name})",
    "note": "L2 holds what the index / covariance kernels of the same step left, so the DRAM traffic can be below the algorithmic bytes",
}
(ROOT / "profiles" / "align_traffic.json").write_text(json.dumps(doc, indent=1) + "\n")
print(json.dumps(doc, indent=1))
