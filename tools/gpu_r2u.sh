#!/bin/bash
# N-GPU run: multi-GPU pytest (N >= 2), N-rank parity check, then the N-rank bench (C3 strong + C5 strong)
TAG=${1:-r2u}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
if [ "${3:-test}" = test ]; then
  timeout 1500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 700 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|E  )" "$OUT/pytest.log" | head; tail -2 "$OUT/pytest.log"
fi
ISFM_SPLIT_MATVEC=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_check.py > "$OUT/check_n$N.log" 2>&1; echo "check exit $?"; grep MULTIGPU_OK "$OUT/check_n$N.log" || tail -5 "$OUT/check_n$N.log"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 10 --warmup 3 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench n$N exit $?"; tail -3 "$OUT/bench_n$N.err"
python - "$OUT" $N <<'P'
import json, sys, os
for f in ("bench_n" + sys.argv[2],):
    try:
        d = json.load(open(os.path.join(sys.argv[1], f + ".json"))); w = d["work"]
        print(f, "ms/step %.3f its/step %.1f us/it %.1f" % (d["ms_per_step"], w["pcg_iters_per_step"], w["us_per_pcg_iter"]), {k: round(v, 1) for k, v in w["pcg_phase_us_per_iter"].items()},
              "excl_pcg %.3f comm %.3f" % (w["ms_per_trial_excl_pcg"], w["comm_ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, d.get("final_robust_cost"), d.get("matvec_split"), d.get("pcg_exchange"))
        print("   e2e", {k: v for k, v in d["e2e"].items() if k in ("value", "seconds", "cold_seconds", "setup_seconds")})
        c5 = d.get("c5")
        if c5:
            print("  c5", {k: c5.get(k) for k in ("ms_per_step", "pcg_iters", "final_robust_cost", "speedup_vs_n1", "cost_rel_diff_vs_n1", "loss_rel_diff_vs_n1_max", "rejects")})
            print("  c5 work", c5["work"]); print("  c5 kernels", {k: round(v["ms_per_step"], 3) for k, v in c5["kernels"].items()})
    except Exception as e:
        print(f, "no line", e)
P
