#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_c5_strong.sh <tag> <N>': strong scaling of C5 (60 M observations, street geometry)
TAG=${1:-c5}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus $N --config C5 --scaling strong --steps 10 --warmup 3 --no-cpu --quick > "$OUT/bench_c5.json" 2> "$OUT/bench_c5.err"; echo "bench exit $?"
tail -4 "$OUT/bench_c5.err"
python - "$OUT/bench_c5.json" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d.get(k) for k in ("value", "ms_per_step", "pcg_iters", "pcg_exchange", "matvec_split", "n_gpus", "final_robust_cost", "rejects")})
    print({k: (round(v["ms_per_step"], 3), round(v["us_per_launch"], 1)) for k, v in d["kernels"].items()})
except Exception as e:
    print("no bench line", e)
P
