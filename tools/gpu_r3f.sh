#!/bin/bash
# last sanity: GP tests (C4 x 0.1 fp32 trajectory), smoke(), C4 line
OUT=gpurun_out/${1:-r3f}; mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gp_gpu.py -q --timeout 200 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|E  )" "$OUT/pytest.log" | head; tail -2 "$OUT/pytest.log"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -3 "$OUT/smoke.log"
timeout 200 python bench.py --config C4 --steps 10 --warmup 3 --no-cpu > "$OUT/c4.json" 2> "$OUT/c4.err"; echo "c4 exit $?"
python - "$OUT" <<'P'
import json, sys, os
d = json.load(open(os.path.join(sys.argv[1], "c4.json"))); print("c4", d["ms_per_step"], d["pcg_iters"], d["losses"][-1], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
P
