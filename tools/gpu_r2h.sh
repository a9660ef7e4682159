#!/bin/bash
# N = 1: tests; C3 / C5 with unit size 48 vs 24; ncu launch list
TAG=${1:-r2h}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest.log"
tail -6 "$OUT/pytest.log"
run() { env $2 timeout 900 python bench.py $3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
run c3_default "ISFM_X=1" "--steps 10 --warmup 3"
run c3_c24 "ISFM_LIB_PATH=$PWD/instantsfm_b200/lib_c24/libisfm_b200.so" "--steps 10 --warmup 3"
run c5_default "ISFM_X=1" "--config C5 --steps 10 --warmup 2"
run c5_c24 "ISFM_LIB_PATH=$PWD/instantsfm_b200/lib_c24/libisfm_b200.so" "--config C5 --steps 10 --warmup 2"
python - "$OUT" <<'P'
import json, sys, os
for f in ("c3_default", "c3_c24", "c5_default", "c5_c24"):
    try:
        d = json.load(open(os.path.join(sys.argv[1], f + ".json"))); w = d["work"]
        print(f, "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()},
              "excl_pcg %.3f" % w["ms_per_trial_excl_pcg"], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if k in ("linearize", "coarse", "pcg_solve")}, d["rejects"])
    except Exception as e:
        print(f, "no line", e)
P
