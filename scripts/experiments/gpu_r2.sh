#!/bin/bash
# Round-2 GPU-box call: parity tests, extra peaks, the default bench line (all workloads), optionally launch list / ncu.
# usage: scripts/gpu_r2.sh <tag> [tests] [peaks] [bench] [launches] [ref]
TAG=${1:-run}; shift; WHAT=" ${*:-tests peaks bench} "
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [[ $WHAT == *" tests "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -80 > $OUT/pytest_gpu.txt
  echo "== pytest"; tail -40 $OUT/pytest_gpu.txt
fi
if [[ $WHAT == *" peaks "* ]]; then
  timeout 300 python scripts/measure_peaks.py $OUT/peaks_extra.json > $OUT/peaks.log 2>&1; cat $OUT/peaks.log | tail -3
  [ -s $OUT/peaks_extra.json ] && cp $OUT/peaks_extra.json profiles/peaks_extra.json
fi
if [[ $WHAT == *" bench "* ]]; then
  timeout 1200 python bench.py --steps 10 --warmup 3 > $OUT/bench_all.json 2> $OUT/bench_all.err
  echo "== bench rc=$?"; cut -c1-6000 $OUT/bench_all.json; tail -5 $OUT/bench_all.err
fi
if [[ $WHAT == *" ref "* ]]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
  echo "== ref rc=$?"; cut -c1-1500 $OUT/bench_ref.json
fi
if [[ $WHAT == *" launches "* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_default.csv python bench.py --steps 2 --warmup 1 --sub-steps 1 --no-cpu --no-e2e > $OUT/ncu_launch.log 2>&1
  python profiles/launch_summary.py $OUT/launches_default.csv > $OUT/launches_default_summary.txt 2>&1; head -30 $OUT/launches_default_summary.txt
fi
