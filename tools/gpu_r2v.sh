#!/bin/bash
# quick N=1 check: solver tests + C3 / C5 / C2 bench lines (+ C4 GP)
OUT=gpurun_out/${1:-r2v}; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py tests/test_ba_gpu.py tests/test_gp_gpu.py tests/test_edge_cases_gpu.py tests/test_processors_gpu.py -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|E  )" "$OUT/pytest.log" | head -20; tail -2 "$OUT/pytest.log"
run() { env $2 timeout 600 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
run c3 "X=1" "--steps 10 --warmup 3"
run c2 "X=1" "--config C2 --steps 10 --warmup 3"
run c5_n1 "X=1" "--config C5 --steps 10 --warmup 2"
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu > "$OUT/c4.json" 2> "$OUT/c4.err"; echo "c4 exit $?"
python - "$OUT" <<'P'
import json, sys, os, glob
for f in sorted(glob.glob(os.path.join(sys.argv[1], "c[235]*.json"))):
    try:
        d = json.load(open(f)); w = d["work"]
        print(os.path.basename(f), "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()}, {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, d["final_robust_cost"], d["rejects"], d["pcg_iters"])
    except Exception as e:
        print(f, "no line", e)
d = json.load(open(os.path.join(sys.argv[1], "c4.json"))); print("c4", d["ms_per_step"], d["pcg_iters"], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
P
