#!/bin/bash
OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"basis_kernel" -s 2 -c 1 -o $OUT/prof_basis python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_basis.log 2>&1
ncu -i $OUT/prof_basis.ncu-rep --page source --csv > $OUT/prof_basis_source.csv 2>/dev/null
python profiles/ncu_source_top.py $OUT/prof_basis_source.csv 45 | cut -c1-210
