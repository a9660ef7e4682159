#!/bin/bash
# N = 1 reference runs for the scaling comparison: GPU tests, C5 (10 steps) fixture, C3
TAG=${1:-r2j}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest.log"
tail -4 "$OUT/pytest.log"
run() { env $2 timeout 900 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
run c3_n1 "ISFM_X=1" "--steps 10 --warmup 3"
run c5_n1 "ISFM_X=1" "--config C5 --steps 10 --warmup 2"
python - "$OUT" <<'P'
import json, sys, os
for f in ("c3_n1", "c5_n1"):
    d = json.load(open(os.path.join(sys.argv[1], f + ".json"))); w = d["work"]
    print(f, "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()},
          "excl_pcg %.3f" % w["ms_per_trial_excl_pcg"], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, d["rejects"], d["final_robust_cost"])
P
