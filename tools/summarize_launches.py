"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches.txt"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
rows = [r for r in rows if r[0] >= skip]
agg = OrderedDict()
for _, name, ns, grid, block in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    if len(short) > 70:
        short = short[:67] + "..."
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
print("launches: %d  total device time: %.3f ms (ncu serialised, cold-cache: compare SHARES)" % (len(rows), total / 1e6))
print("%-72s %7s %12s %8s %10s" % ("kernel", "count", "total ms", "share", "avg us"))
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %7d %12.3f %7.2f%% %10.1f" % (k, c, ns / 1e6, 100 * ns / total, ns / c / 1e3))
