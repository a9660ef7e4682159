#!/bin/bash
# Final single-GPU evidence run of a round: tests, default bench, launch list, ncu capture of K1, GP bench.
OUT=gpurun_out/${1:-final}; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -x -q > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; tail -4 "$OUT/pytest.log"
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"
timeout 200 python tools/gp_bench.py > "$OUT/gp_bench.json" 2> "$OUT/gp_bench.err"; echo "gp exit $?"; tail -2 "$OUT/gp_bench.json" | cut -c1-600
export ISFM_NO_GRAPH=1
timeout 300 python tools/prof_run.py > "$OUT/prof_run.log" 2>&1; PR=$?; echo "prof_run exit $PR"
if [ $PR -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
    python tools/prof_run.py > "$OUT/ncu_launch.log" 2>&1; echo "launch list exit $?"
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:fused_linearize_tma" -s 1 -c 1 -f -o "$OUT/full_fused_linearize_tma" \
    python tools/prof_run.py > "$OUT/ncu_k1.log" 2>&1; echo "ncu k1 exit $?"
fi
