#!/bin/bash
# new stage-stream loop: correctness tests, C3 bench; two-level on dense systems: C1 parity, C3 bench
OUT=gpurun_out/${1:-r2o}; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py tests/test_ba_gpu.py tests/test_gp_gpu.py -q --timeout 600 -k "not c1_full" > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; tail -4 "$OUT/pytest.log"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"
ISFM_TWO_LEVEL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick > "$OUT/bench_2l.json" 2> "$OUT/bench_2l.err"; echo "bench 2l exit $?"
ISFM_TWO_LEVEL=1 ISFM_COARSE_MAX_CLUSTERS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick > "$OUT/bench_2l1.json" 2> "$OUT/bench_2l1.err"; echo "bench 2l1 exit $?"
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu > "$OUT/bench_c4.json" 2> "$OUT/bench_c4.err"; echo "bench c4 exit $?"
python - "$OUT" <<'P'
import json, sys, os
for f in ("bench", "bench_2l", "bench_2l1"):
    try:
        d = json.load(open(os.path.join(sys.argv[1], f + ".json"))); w = d["work"]
        print(f, "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()}, {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, d["final_robust_cost"], d["rejects"])
    except Exception as e:
        print(f, "no line", e)
d = json.load(open(os.path.join(sys.argv[1], "bench_c4.json"))); print("c4", d["ms_per_step"], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
P
for v in "ISFM_TWO_LEVEL=1" "ISFM_TWO_LEVEL=1 ISFM_COARSE_MAX_CLUSTERS=1"; do
  echo "== $v" | tee -a "$OUT/parity.log"
  env $v timeout 300 python tools/parity_diag2.py c1 1e-6,1e-7 2>&1 | grep -v Warn | tee -a "$OUT/parity.log" | cut -c1-700
  env $v timeout 300 python tools/parity_diag2.py proc 1e-6 2>&1 | grep -v Warn | tee -a "$OUT/parity.log" | cut -c1-500
done
