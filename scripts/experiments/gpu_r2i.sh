#!/bin/bash
TAG=${1:-R2i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python scripts/experiments/repro_sharded.py 2 200000 > $OUT/repro.txt 2>&1; echo "rc=$?"; grep -v "^frame" $OUT/repro.txt | tail -25
timeout 1500 compute-sanitizer --tool memcheck --print-limit 6 python scripts/experiments/repro_sharded.py 2 20000 > $OUT/memcheck.txt 2>&1; echo "memcheck rc=$?"
grep -B2 -A16 "Invalid\|ERROR SUMMARY" $OUT/memcheck.txt | head -90
python scripts/level1_latency.py 2>&1 | tail -3 | tee $OUT/level1_latency.txt
