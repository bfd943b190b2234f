"""Instruction-count hot spots from an .ncu-rep source page: python profiles/sass_hot.py rep kernel_regex"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ie, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = [(r[ia].strip(), int(r[ie]), int(r[ist])) for r in rows[2:] if len(r) > ist]
tot = sum(d[1] for d in data)
print("total instructions", tot, "stall samples", sum(d[2] for d in data))
i = 0
while i < len(data):
    j, s, st = i, 0, 0
    while j < len(data) and abs(data[j][1] - data[i][1]) <= 0.2 * max(data[i][1], 1):
        s += data[j][1]
        st += data[j][2]
        j += 1
    ops = " ".join(d[0].split()[0] if not d[0].startswith("@") else d[0].split()[1] for d in data[i:j])
    if s / tot > 0.004:
        print(f"[{i:4d}-{j - 1:4d}] n={j - i:3d} each~{data[i][1]:>9d} inst={s / tot * 100:5.1f}% stall={st:6d}  {ops[:120]}")
    i = j
