"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name.
Usage: python tools/launch_summary.py gpurun_out/launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[ui]
    v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    agg[name].append(v)
tot = sum(sum(v) for v in agg.values())
print("%-70s %6s %10s %9s %9s %6s" % ("kernel", "n", "total us", "avg us", "min us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-70s %6d %10.1f %9.1f %9.1f %5.1f%%" % (k[:70], len(v), sum(v), sum(v) / len(v), min(v), 100 * sum(v) / tot))
