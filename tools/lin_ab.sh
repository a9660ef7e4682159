#!/bin/bash
# fused linearize kernel on one GPU: GPU tests + per-kernel CUDA-event time from bench.py's profile pass.
OUT=gpurun_out/${1:-linab}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo pytest exit $?; tail -4 $OUT/pytest.log
run() { env $2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > $OUT/bench_$1.json 2>/dev/null; python -c "
import json,sys; d=json.load(open('$OUT/bench_$1.json')); k=d['kernels']['linearize']; print('$1', 'linearize %.1f us %.0f GB/s' % (k['us_per_launch'], k['achieved_gbs']), 'step %.3f ms' % d['ms_per_step'], d['final_robust_cost'])"; }
run tma ISFM_X=1
