"""Summarise an `ncu --metrics gpu__time_duration.sum[,smsp__inst_executed.sum] --csv` launch list."""
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
