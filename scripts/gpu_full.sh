#!/bin/bash
# full single-GPU validation: GPU test suite, smoke, the driver's default line and reference arm
TAG=${1:-R2u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > $OUT/pytest_gpu.txt 2>&1; tail -5 $OUT/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.txt 2>&1; tail -2 $OUT/smoke.txt
bash scripts/gpu_default_n.sh $TAG 1
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err; cut -c1-300 $OUT/bench_reference_n1.json
