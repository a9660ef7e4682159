#!/bin/bash
# 16-warp / 3-stage stream: correctness tests, C3 bench; two-level cluster-count sweep on dense C3; C5 N=1
OUT=gpurun_out/${1:-r2p}; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py tests/test_ba_gpu.py tests/test_gp_gpu.py tests/test_edge_cases_gpu.py -q --timeout 600 -k "not c1_full" > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; tail -4 "$OUT/pytest.log"
run() { env $2 timeout 600 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
run bench "X=1" "--steps 10 --warmup 3"
for c in 2 4 16 32; do run bench_2l_$c "ISFM_TWO_LEVEL=1 ISFM_COARSE_MAX_CLUSTERS=$c" "--steps 10 --warmup 3"; done
run bench_2l "ISFM_TWO_LEVEL=1" "--steps 10 --warmup 3"
run c2_2l "ISFM_TWO_LEVEL=1" "--config C2 --steps 10 --warmup 3"
run c2 "X=1" "--config C2 --steps 10 --warmup 3"
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
