#!/bin/bash
# parity diagnostics under solver variants + C3 bench with the per-iteration-kernel PCG path
OUT=gpurun_out/${1:-r2n}; mkdir -p "$OUT"
for v in "X=1" "ISFM_PCG_VERIFY=1" "ISFM_NO_PERSISTENT=1" "ISFM_NO_TMA=1" "ISFM_NO_FUSED=1"; do
  echo "== $v" | tee -a "$OUT/parity.log"
  env $v timeout 300 python tools/parity_diag2.py c1 1e-6,1e-7 2>&1 | grep -v Warn | tee -a "$OUT/parity.log"
  env $v timeout 300 python tools/parity_diag2.py proc 1e-6,1e-7 2>&1 | grep -v Warn | tee -a "$OUT/parity.log"
done
ISFM_NO_PERSISTENT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick > "$OUT/bench_nopersist.json" 2> "$OUT/bench_nopersist.err"; echo "bench exit $?"
python - "$OUT" <<'P'
import json, sys, os
d = json.load(open(os.path.join(sys.argv[1], "bench_nopersist.json"))); w = d["work"]
print("nopersist ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
P
