#!/bin/bash
OUT=gpurun_out/${1:-dbg}; mkdir -p $OUT
for i in 1 2 3 4 5 6; do
  timeout 120 python scripts/apply_micro.py --mmax 200 --ann 4096 2>&1 | tail -12
done > $OUT/rep.txt 2>&1
tail -40 $OUT/rep.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/apply_micro.py --mmax 200 --ann 4096 --nq 1024 --reps 1 > $OUT/memcheck.txt 2>&1
tail -30 $OUT/memcheck.txt
