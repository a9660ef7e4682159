#!/bin/bash
# Round-2 single-GPU run: GPU tests, default bench, C5 at N = 1, combine-phase variants.
# Usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2b.sh r2b'
TAG=${1:-r2b}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest.log"
tail -8 "$OUT/pytest.log"
timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"; tail -3 "$OUT/bench.err"
timeout 900 python bench.py --config C5 --steps 5 --warmup 2 --no-cpu --quick > "$OUT/c5_n1.json" 2> "$OUT/c5_n1.err"; echo "c5 exit $?"; tail -3 "$OUT/c5_n1.err"
for w in 1 2 4; do
  ISFM_PCG_WPR=$w timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick > "$OUT/bench_wpr$w.json" 2> "$OUT/bench_wpr$w.err"; echo "wpr $w exit $?"
done
python - "$OUT" <<'P'
import json, sys, os
for f in ("bench", "c5_n1", "bench_wpr1", "bench_wpr2", "bench_wpr4"):
    try:
        d = json.load(open(os.path.join(sys.argv[1], f + ".json")))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "pcg_iters", "final_robust_cost", "rejects")})
        print("   work", d.get("work"))
        print("   kernels", {k: (round(v["ms_per_step"], 3), round(v["us_per_launch"], 1), v.get("frac_algorithmic") and round(v["frac_algorithmic"], 3)) for k, v in d["kernels"].items()})
        for k in ("e2e", "e2e_dropin", "reference_gpu"):
            if d.get(k): print("   ", k, {kk: vv for kk, vv in d[k].items() if kk not in ("costs", "what", "includes", "note")})
        for k in ("cpu_baseline", "c1"):
            if d.get(k): print("   ", k, d[k].get("ms_per_lm_step"), d[k].get("gpu_same_problem"), d[k].get("parity"))
    except Exception as e:
        print(f, "no line", e)
P
