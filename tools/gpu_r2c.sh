#!/bin/bash
# parity diagnostic, K1 occupancy variants, tests, C5
TAG=${1:-r2c}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python tools/parity_diag.py 2>&1 | grep -v Warning | tee "$OUT/parity.log" | grep "tol 1e-06\|tol 3e-07"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest.log"
tail -6 "$OUT/pytest.log"
for v in "" _m4 _m3; do
  ISFM_LIB_PATH=$PWD/instantsfm_b200/lib$v/libisfm_b200.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick > "$OUT/bench$v.json" 2> "$OUT/bench$v.err"; echo "bench$v exit $?"
done
timeout 900 python bench.py --config C5 --steps 5 --warmup 2 --no-cpu --quick > "$OUT/c5_n1.json" 2> "$OUT/c5_n1.err"; echo "c5 exit $?"; tail -3 "$OUT/c5_n1.err"
python - "$OUT" <<'P'
import json, sys, os
for f in ("bench", "bench_m4", "bench_m3", "c5_n1"):
    try:
        d = json.load(open(os.path.join(sys.argv[1], f + ".json")))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "pcg_iters", "final_robust_cost", "rejects")})
        print("   work", d.get("work"))
        print("   kernels", {k: (round(v["ms_per_step"], 3), round(v["us_per_launch"], 1), v.get("frac_algorithmic") and round(v["frac_algorithmic"], 3)) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "no line", e)
P
