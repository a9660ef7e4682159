#!/bin/bash
# ridge x damping-floor sweep of the two-level preconditioner on dense C3 (25 LM steps) and C1
OUT=gpurun_out/${1:-r2s}; mkdir -p "$OUT"
run() { env $2 timeout 600 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
for r in 0 4 32; do for m in 0 1e-7 1e-6 1e-5; do
  run c3_r${r}_m$m "ISFM_COARSE_RIDGE=$r ISFM_MIN_DAMPING=$m" "--steps 25 --warmup 3"
done; done
for m in 1e-6 1e-5; do run c3_bj_m$m "ISFM_TWO_LEVEL=0 ISFM_MIN_DAMPING=$m" "--steps 25 --warmup 3"; done
for r in 0 4; do for c in 111 444 55; do run c3_r${r}_m1e-6_cs$c "ISFM_COARSE_RIDGE=$r ISFM_MIN_DAMPING=1e-6 ISFM_COARSE_CS=$c" "--steps 25 --warmup 3"; done; done
python - "$OUT" <<'P'
import json, sys, os, glob
for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.json"))):
    try:
        d = json.load(open(f)); w = d["work"]
        print(os.path.basename(f), "ms/step %.3f its/step %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if k in ("pcg_solve", "coarse")}, "%.3f" % d["final_robust_cost"], d["rejects"], d["pcg_iters"])
    except Exception as e:
        print(f, "no line", e)
P
