#!/bin/bash
# A/B of the mat-vec stream order / contiguous-p path on one box, plus the graph path as reference
TAG=${1:-r2d}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py tests/test_ba_gpu.py -m gpu -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest.log"
tail -4 "$OUT/pytest.log"
run() { # name env config-args
  env $2 timeout 600 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"
}
run c3_default "ISFM_X=1" "--steps 10 --warmup 3"
run c3_blocked "ISFM_PCG_BLOCKED_ORDER=1" "--steps 10 --warmup 3"
run c3_nocontig "ISFM_PCG_NO_CONTIG=1" "--steps 10 --warmup 3"
run c3_graph "ISFM_NO_PERSISTENT=1" "--steps 10 --warmup 3"
run c5_default "ISFM_X=1" "--config C5 --steps 5 --warmup 2"
run c5_blocked "ISFM_PCG_BLOCKED_ORDER=1" "--config C5 --steps 5 --warmup 2"
run c2_default "ISFM_X=1" "--config C2 --steps 10 --warmup 3"
python - "$OUT" <<'P'
import json, sys, os
for f in ("c3_default", "c3_blocked", "c3_nocontig", "c3_graph", "c5_default", "c5_blocked", "c2_default"):
    try:
        d = json.load(open(os.path.join(sys.argv[1], f + ".json")))
        w = d["work"]
        print(f, "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()},
              "excl_pcg %.3f" % w["ms_per_trial_excl_pcg"], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if k in ("linearize", "coarse", "pcg_solve", "pcg_vec")})
    except Exception as e:
        print(f, "no line", e)
P
