#!/bin/bash
# per-kernel durations of the coarse-level kernels (C5, 2 LM steps)
TAG=${1:-r2i}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python tools/prof_run.py --config C5 --steps 2 > "$OUT/prof_run.log" 2>&1; echo "prof_run exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:coarse|gj_" -c 400 --csv --log-file "$OUT/coarse_launches.csv" python tools/prof_run.py --config C5 --steps 2 > "$OUT/ncu.log" 2>&1; echo "ncu exit $?"
python - "$OUT/coarse_launches.csv" <<'P'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)   # -> us
    k = r[ki].split("(")[0][:60]
    agg[k][0] += 1; agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n %4d total %9.1f us  avg %8.1f us" % (k, n, t, t / n))
P
