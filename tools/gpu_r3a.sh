#!/bin/bash
# N-rank C5 window (LM steps 4-6) under variants
N=${2:-8}; OUT=gpurun_out/${1:-r3a}; mkdir -p "$OUT"
run() { env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --config C5 --steps 3 --warmup 3 --no-cpu --quick > "$OUT/$1.json" 2> "$OUT/$1.err"; echo "$1 exit $?"; }
run c5_def "X=1"
run c5_u16 "ISFM_UNITS_PER_WARP=16"
run c5_u32 "ISFM_UNITS_PER_WARP=32"
run c5_grid "ISFM_PEER_PUSH=grid"
python - "$OUT" <<'P'
import json, sys, os, glob
for f in sorted(glob.glob(os.path.join(sys.argv[1], "*.json"))):
    try:
        d = json.load(open(f)); w = d["work"]
        print(os.path.basename(f), "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()}, "excl %.3f comm %.3f" % (w["ms_per_trial_excl_pcg"], w["comm_ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, d["final_robust_cost"], d["rejects"], d["pcg_iters"])
    except Exception as e:
        print(f, "no line", e)
P
