#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> <N> [test]': multi-rank parity check (both transports)
# and the N-rank bench with the peer-memory exchange and with NCCL.
TAG=${1:-multi}; N=${2:-2}; TEST=${3:-test}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
if [ "$TEST" = test ]; then
  timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; tail -15 "$OUT/pytest.log"
fi
run() { # name, extra env
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu $3 > "$OUT/bench_$1.json" 2> "$OUT/bench_$1.err"; echo "bench $1 exit $?"
  python - "$OUT/bench_$1.json" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d.get(k) for k in ("value", "ms_per_step", "pcg_iters_per_step", "pcg_exchange", "matvec_split", "n_gpus", "scaling", "final_robust_cost")})
    print({k: (round(v["ms_per_step"], 3), round(v["us_per_launch"], 1)) for k, v in d["kernels"].items()})
except Exception as e:
    print("no bench line", e)
P
}
if [ "${4:-}" != nobench ]; then
run split "ISFM_X=1" ""
run nosplit "ISFM_SPLIT_MATVEC=0" ""
fi
if [ "${4:-}" = nccl ]; then run nccl "ISFM_NO_PEER=1 ISFM_SPLIT_MATVEC=0" ""; fi
ls "$OUT"
