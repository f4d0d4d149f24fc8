#!/bin/bash
TAG=${1:-R2j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 compute-sanitizer --tool memcheck --print-limit 6 python scripts/experiments/repro_sharded.py 2 20000 > $OUT/memcheck.txt 2>&1; echo "memcheck rc=$?"
grep -v "^=========     \|^frame" $OUT/memcheck.txt | tail -12
timeout 900 python scripts/experiments/repro_sharded.py 2 200000 > $OUT/repro.txt 2>&1; echo "rc=$?"; grep -v "^frame" $OUT/repro.txt | tail -8
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
