#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, sanitizer on a subset, bench, ncu captures.
# usage: scripts/gpu_check.sh <tag> [tests|bench|ncu|all]
TAG=${1:-run}; WHAT=${2:-all}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [[ $WHAT == all || $WHAT == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $OUT/pytest.txt
  tail -5 $OUT/pytest.txt
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "basis_and_dense_counts_random or golden_rule_fixtures and miqs or long_sequences" > $OUT/memcheck.txt 2>&1
  echo "memcheck exit $?"; tail -4 $OUT/memcheck.txt
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
  cat $OUT/bench.json; tail -3 $OUT/bench.err
fi
if [[ $WHAT == all || $WHAT == ncu ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"count_dense_kernel|basis_kernel" -s 2 -c 4 -o $OUT/prof python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_full.log 2>&1
  ls -la $OUT
fi
