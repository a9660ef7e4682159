#!/bin/bash
# C3 quick bench (persistent vs graph on the same box) + ncu --set full of the persistent PCG kernel
TAG=${1:-r2e}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
run() { env $2 timeout 600 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
run c3_default "ISFM_X=1" "--steps 10 --warmup 3"
run c3_graph "ISFM_NO_PERSISTENT=1" "--steps 10 --warmup 3"
python - "$OUT" <<'P'
import json, sys, os
for f in ("c3_default", "c3_graph"):
    d = json.load(open(os.path.join(sys.argv[1], f + ".json"))); w = d["work"]
    print(f, "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()},
          {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if k in ("linearize", "pcg_solve", "pcg_vec")})
P
timeout 300 python tools/prof_run.py --steps 3 > "$OUT/prof_run.log" 2>&1; echo "prof_run exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pcg_persistent" -s 2 -c 1 -f -o "$OUT/full_pcg_persistent" python tools/prof_run.py --steps 3 > "$OUT/ncu_pcg.log" 2>&1; echo "ncu exit $?"; tail -3 "$OUT/ncu_pcg.log"
ls -la "$OUT"
