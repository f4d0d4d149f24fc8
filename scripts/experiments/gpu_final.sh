#!/bin/bash
# round-end evidence on one B200: smoke, full GPU parity suite, every bench workload, reference arm, ncu launch list of the default bench
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 > $OUT/pytest_gpu.txt; tail -3 $OUT/pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "== reference rc=$?"; cut -c1-200 $OUT/bench_reference.json
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_vectorize.json 2> $OUT/bench_vectorize.err; echo "== vectorize rc=$?"; cut -c1-1700 $OUT/bench_vectorize.json
for w in apply learn apply_sparse sweep; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "== $w rc=$?"; cut -c1-330 $OUT/bench_$w.json; tail -2 $OUT/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_vectorize.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_vectorize.log 2>&1
python profiles/launch_summary.py $OUT/launches_vectorize.csv 12
