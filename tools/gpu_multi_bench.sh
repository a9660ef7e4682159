#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_multi_bench.sh <tag> <N> [extra env]': one N-rank bench run (default settings).
TAG=${1:-multi}; N=${2:-2}; EXTRA=${3:-ISFM_X=1}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
env $EXTRA timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"
tail -5 "$OUT/bench.err"
python - "$OUT/bench.json" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d.get(k) for k in ("value", "ms_per_step", "pcg_iters_per_step", "pcg_iters", "pcg_exchange", "matvec_split", "n_gpus", "final_robust_cost")})
    print({k: (round(v["ms_per_step"], 3), round(v["us_per_launch"], 1)) for k, v in d["kernels"].items()})
    print(d["e2e"])
except Exception as e:
    print("no bench line", e)
P
