#!/bin/bash
OUT=gpurun_out/r1k; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "not tensor_cores" 2>&1 | tail -30 > $OUT/pytest_main.txt; tail -6 $OUT/pytest_main.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -k "tensor_cores" 2>&1 | tail -30 > $OUT/pytest_tc.txt; tail -4 $OUT/pytest_tc.txt
for w in learn sweep; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "== $w rc=$?"; cut -c1-300 $OUT/bench_$w.json; tail -3 $OUT/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_sweep.csv python bench.py --workload sweep --nseq 200000 --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_sweep.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_learn.csv python bench.py --workload learn --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_learn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_sort_kernel" -s 4 -c 3 -o $OUT/prof_csrsort python bench.py --workload sweep --nseq 200000 --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_full_csrsort.log 2>&1
ncu -i $OUT/prof_csrsort.ncu-rep --page raw --csv > $OUT/prof_csrsort_raw.csv 2>/dev/null
ls -la $OUT | head -30
