#!/bin/bash
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"apply_sparse_kernel" -s 2 -c 1 -o $OUT/prof_as python bench.py --workload apply_sparse --nseq 50000 --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_as.log 2>&1
ncu -i $OUT/prof_as.ncu-rep --page raw --csv > $OUT/prof_as_raw.csv 2>/dev/null
ncu -i $OUT/prof_as.ncu-rep --page source --csv > $OUT/prof_as_source.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_as_raw.csv | cut -c1-150
python profiles/ncu_source_top.py $OUT/prof_as_source.csv 30 | cut -c1-200
