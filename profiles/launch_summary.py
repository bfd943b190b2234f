"""Summarise an `ncu --metrics gpu__time_duration.sum[,smsp__inst_executed.sum] --csv` launch list.

    python profiles/launch_summary.py launches.csv [units]     units: registrations the run made (adds per-registration totals)"""
import csv
import sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = defaultdict(lambda: defaultdict(list))
for row in csv.DictReader(lines):
    agg[row["Kernel Name"][:56]][row["Metric Name"]].append(float(row["Metric Value"].replace(",", "")))
tot = sum(sum(v["gpu__time_duration.sum"]) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = v["gpu__time_duration.sum"]
    i = v.get("smsp__inst_executed.sum")
    extra = f" inst mean={sum(i) / len(i) / 1e6:8.2f}M" if i else ""
    print(f"{k:56s} n={len(t):3d} total={sum(t) / 1e3:9.1f}us ({100 * sum(t) / tot:4.1f}%) mean={sum(t) / len(t) / 1e3:8.1f}us min={min(t) / 1e3:8.1f}{extra}")

units = int(sys.argv[2]) if len(sys.argv) > 2 else 0
inst_tot = sum(sum(v.get("smsp__inst_executed.sum", [])) for v in agg.values())
print(f"total: {tot / 1e3:.1f} us of kernel time, {inst_tot / 1e6:.1f} M warp instructions over {sum(len(v['gpu__time_duration.sum']) for v in agg.values())} launches")
if units:
    print(f"per registration ({units} units): {tot / 1e3 / units:.1f} us of (serialised) kernel time, {inst_tot / 1e6 / units:.2f} M warp instructions")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("smsp__inst_executed.sum", [0]))):
        i = v.get("smsp__inst_executed.sum")
        if i:
            print(f"  {k:56s} {sum(i) / 1e6 / units:8.2f} M inst / registration  ({100 * sum(i) / max(inst_tot, 1):4.1f}%)")
