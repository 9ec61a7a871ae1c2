"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
print("total kernel time %.3f ms over %d launches" % (total / 1e6, sum(v[0] for v in tot.values())))
print("%-70s %6s %10s %7s" % ("kernel", "count", "ms", "share"))
for k, (c, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %6d %10.3f %6.1f%%" % (k[:70], c, ns / 1e6, 100 * ns / total))
