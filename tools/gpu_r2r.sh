#!/bin/bash
# two-level by default (ridge, fp64 inverse, fallback): full GPU tests, C3 x 25 steps stress, C1/C2/C3/C5 bench lines
OUT=gpurun_out/${1:-r2r}; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|E  )" "$OUT/pytest.log" | head -20; tail -2 "$OUT/pytest.log"
run() { env $2 timeout 600 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
run c3 "X=1" "--steps 10 --warmup 3"
run c3_long "X=1" "--steps 25 --warmup 3"
run c3_long_cs111 "ISFM_COARSE_CS=111" "--steps 25 --warmup 3"
run c3_long_cs444 "ISFM_COARSE_CS=444" "--steps 25 --warmup 3"
run c3_long_bj "ISFM_TWO_LEVEL=0" "--steps 25 --warmup 3"
run c2 "X=1" "--config C2 --steps 10 --warmup 3"
run c1 "X=1" "--config C1 --steps 10 --warmup 3"
run c5_n1 "X=1" "--config C5 --steps 10 --warmup 2"
python - "$OUT" <<'P'
import json, sys, os, glob
for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.json"))):
    try:
        d = json.load(open(f)); w = d["work"]
        print(os.path.basename(f), "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()}, {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if k in ("pcg_solve", "coarse", "misc")}, d["final_robust_cost"], d["rejects"], d["pcg_iters"])
    except Exception as e:
        print(f, "no line", e)
P
