#!/bin/bash
# All bench workloads on one GPU. usage: scripts/gpu_bench_all.sh <tag>
TAG=${1:-run}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for w in vectorize apply learn apply_sparse; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  echo "== $w rc=$?"; cut -c1-1500 $OUT/bench_$w.json; tail -5 $OUT/bench_$w.err
done
