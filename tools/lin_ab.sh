#!/bin/bash
# A/B of the fused linearize kernel variants on one GPU: per-kernel CUDA-event time from bench.py's profile pass.
OUT=gpurun_out/${1:-linab}; mkdir -p $OUT
run() { env $2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > $OUT/bench_$1.json 2>/dev/null; python -c "
import json,sys; d=json.load(open('$OUT/bench_$1.json')); k=d['kernels']['linearize']; print('$1', 'linearize %.1f us %.0f GB/s' % (k['us_per_launch'], k['achieved_gbs']), 'step %.3f ms' % d['ms_per_step'], d['final_robust_cost'])"; }
run tma32 ISFM_TMA_BOX=32
run tma64 ISFM_TMA_BOX=64

run notma ISFM_NO_TMA=1

