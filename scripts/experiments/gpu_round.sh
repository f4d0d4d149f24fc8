#!/bin/bash
# One GPU-box call: parity tests (tensor-core tests in their own process), then every bench workload.
# usage: scripts/gpu_round.sh <tag>
TAG=${1:-run}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "not tensor_cores" 2>&1 | tail -40 > $OUT/pytest_main.txt
echo "== pytest main"; tail -8 $OUT/pytest_main.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -k "tensor_cores" 2>&1 | tail -60 > $OUT/pytest_tc.txt
echo "== pytest tc"; tail -15 $OUT/pytest_tc.txt
for w in vectorize apply learn apply_sparse; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  echo "== $w rc=$?"; cut -c1-1800 $OUT/bench_$w.json; tail -5 $OUT/bench_$w.err
done
