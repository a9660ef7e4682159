#!/bin/bash
# C5 N=1: robustness of the late LM steps vs damping floor / true-residual verification / PCG tolerance
OUT=gpurun_out/${1:-r2y}; mkdir -p "$OUT"
run() { env $2 timeout 600 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
run c5_def "X=1" "--config C5 --steps 12 --warmup 2"
run c5_m3e-6 "ISFM_MIN_DAMPING=3e-6" "--config C5 --steps 12 --warmup 2"
run c5_m1e-5 "ISFM_MIN_DAMPING=1e-5" "--config C5 --steps 12 --warmup 2"
run c5_verify "ISFM_PCG_VERIFY=1" "--config C5 --steps 12 --warmup 2"
python - "$OUT" <<'P'
import json, sys, os, glob
for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.json"))):
    try:
        d = json.load(open(f)); w = d["work"]
        print(os.path.basename(f), "ms/step %.3f its/step %.1f us/it %.1f trials %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"], w["trials_per_step"]), "%.1f" % d["final_robust_cost"], d["rejects"], d["pcg_iters"], ["%.0f" % x for x in d["losses"]])
    except Exception as e:
        print(f, "no line", e)
P
