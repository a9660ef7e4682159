#!/bin/bash
# GP Schur kernel check + C4 line; C5 N=1 fixture for the scaling window (LM steps 4-6)
OUT=gpurun_out/${1:-r2z}; mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gp_gpu.py tests/test_processors_gpu.py -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|E  )" "$OUT/pytest.log" | head; tail -2 "$OUT/pytest.log"
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu > "$OUT/c4.json" 2> "$OUT/c4.err"; echo "c4 exit $?"
timeout 600 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu --quick > "$OUT/c5_n1.json" 2> "$OUT/c5_n1.err"; echo "c5 exit $?"
timeout 600 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu --quick > "$OUT/c5_n1_b.json" 2> "$OUT/c5_n1_b.err"; echo "c5 b exit $?"
python - "$OUT" <<'P'
import json, sys, os
d = json.load(open(os.path.join(sys.argv[1], "c4.json"))); print("c4", d["ms_per_step"], d["pcg_iters"], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
for f in ("c5_n1", "c5_n1_b"):
    d = json.load(open(os.path.join(sys.argv[1], f + ".json"))); w = d["work"]
    print(f, "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()}, "excl %.3f" % w["ms_per_trial_excl_pcg"], d["final_robust_cost"], d["rejects"], d["pcg_iters"], d["losses"])
P
