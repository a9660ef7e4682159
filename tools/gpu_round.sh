#!/bin/bash
# One gpurun call: GPU tests, the default bench, the ncu launch list and one --set full capture per hot
# kernel (each ncu pass only after the same command exited 0 without ncu).  Outputs under gpurun_out/<tag>/.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1b [tests|notests] [kernel regexes...]'
TAG=${1:-run}; shift
TESTS=${1:-tests}; shift
KERNELS=("$@")
if [ ${#KERNELS[@]} -eq 0 ]; then
  KERNELS=(pcg_spmv_upper schur_offdiag fused_linearize camera_schur backsub_tiles)
fi
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
if [ "$TESTS" = tests ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest.log"
  tail -5 "$OUT/pytest.log"
fi
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"
cat "$OUT/bench.json" | head -c 600; echo
timeout 300 python tools/filter_bench.py > "$OUT/filter_bench.json" 2> "$OUT/filter_bench.err"; echo "filter_bench exit $?"; cat "$OUT/filter_bench.json"
# ncu cannot see kernel nodes inside a graph with conditional nodes: the profiled runs launch the PCG
# iterations as plain stream launches (ISFM_NO_GRAPH=1, same kernels, host polls every 8 iterations)
export ISFM_NO_GRAPH=1
timeout 300 python tools/prof_run.py > "$OUT/prof_run.log" 2>&1; PR=$?; echo "prof_run exit $PR"
if [ $PR -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
    python tools/prof_run.py > "$OUT/ncu_launch.log" 2>&1; echo "launch list exit $?"
  for k in "${KERNELS[@]}"; do
    timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o "$OUT/full_$k" \
      python tools/prof_run.py > "$OUT/ncu_$k.log" 2>&1; echo "ncu $k exit $?"
  done
fi
ls -la "$OUT"
