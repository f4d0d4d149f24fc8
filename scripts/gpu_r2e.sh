#!/bin/bash
OUT=gpurun_out/r2e; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_keys_kernel" -s 0 -c 1 -o $OUT/prof_wk python bench.py --workload learn --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_wk.log 2>&1
ncu -i $OUT/prof_wk.ncu-rep --page raw --csv > $OUT/prof_wk_raw.csv 2>/dev/null
ncu -i $OUT/prof_wk.ncu-rep --page source --csv > $OUT/prof_wk_source.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_wk_raw.csv | cut -c1-150
python profiles/ncu_source_top.py $OUT/prof_wk_source.csv 28 | cut -c1-200
