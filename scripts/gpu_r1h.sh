#!/bin/bash
# confidence + wide + dispatch tests, sweep bench, launch list of the sweep
OUT=gpurun_out/r1h; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_confidence.py tests/test_gpu_wide.py -m gpu -q --timeout 300 2>&1 | tail -40 > $OUT/pytest.txt; tail -15 $OUT/pytest.txt
timeout 900 python bench.py --workload sweep --steps 3 --warmup 3 > $OUT/bench_sweep.json 2> $OUT/bench_sweep.err
echo "== sweep rc=$?"; cut -c1-6000 $OUT/bench_sweep.json; tail -5 $OUT/bench_sweep.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_sweep.csv python bench.py --workload sweep --nseq 100000 --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_sweep.log 2>&1
tail -3 $OUT/ncu_sweep.log
