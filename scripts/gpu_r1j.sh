#!/bin/bash
OUT=gpurun_out/r1j; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_dist_gpu.py -m gpu -q --timeout 300 -x -k "learn or sparse or dist" 2>&1 | tail -40 > $OUT/pytest.txt; tail -15 $OUT/pytest.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "grouped_many or (learn_sparse_matches and 5-3)" > $OUT/memcheck.txt 2>&1; echo memcheck rc=$?; tail -4 $OUT/memcheck.txt
timeout 600 python bench.py --workload learn --steps 5 --warmup 3 > $OUT/bench_learn.json 2> $OUT/bench_learn.err; echo "learn rc=$?"; cut -c1-700 $OUT/bench_learn.json; tail -3 $OUT/bench_learn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_learn.csv python bench.py --workload learn --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_learn.log 2>&1
