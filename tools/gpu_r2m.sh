#!/bin/bash
# Round-2 single-GPU evidence run: GPU tests, default bench (C3, all legs), GP bench (C4), C5 N=1 line,
# ncu launch list and one --set full capture per hot kernel (each ncu pass only after the same
# command exited 0 without ncu).  Outputs under gpurun_out/<tag>/.
TAG=${1:-r2m}; shift
KERNELS=("$@")
if [ ${#KERNELS[@]} -eq 0 ]; then
  KERNELS=(pcg_persistent schur_offdiag fused_linearize_tma camera_schur backsub_tiles)
fi
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest.log"
grep -E "^(FAILED|ERROR|E  )" "$OUT/pytest.log" | head -40; tail -3 "$OUT/pytest.log"
timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"
head -c 400 "$OUT/bench.json"; echo
timeout 300 python bench.py --config C4 --steps 10 --warmup 3 > "$OUT/bench_c4.json" 2> "$OUT/bench_c4.err"; echo "bench C4 exit $?"
head -c 300 "$OUT/bench_c4.json"; echo
timeout 600 python bench.py --config C5 --steps 10 --warmup 2 --no-cpu --quick > "$OUT/c5_n1.json" 2> "$OUT/c5_n1.err"; echo "c5 exit $?"
head -c 300 "$OUT/c5_n1.json"; echo
timeout 300 python tools/prof_run.py > "$OUT/prof_run.log" 2>&1; PR=$?; echo "prof_run exit $PR"
if [ $PR -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
    python tools/prof_run.py > "$OUT/ncu_launch.log" 2>&1; echo "launch list exit $?"
  for k in "${KERNELS[@]}"; do
    timeout 500 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -f -o "$OUT/full_$k" \
      python tools/prof_run.py > "$OUT/ncu_$k.log" 2>&1; echo "ncu $k exit $?"
    # gpurun_out/ is capped at 64 MiB: keep the text pages, drop the report
    ncu -i "$OUT/full_$k.ncu-rep" --page raw --csv > "$OUT/full_$k.raw.csv" 2>/dev/null
    ncu -i "$OUT/full_$k.ncu-rep" --page source --csv --print-source sass 2>/dev/null | gzip > "$OUT/full_$k.source.csv.gz"
    rm -f "$OUT/full_$k.ncu-rep"
  done
fi
ls -la "$OUT"
