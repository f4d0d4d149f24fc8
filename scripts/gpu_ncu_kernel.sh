#!/bin/bash
# `ncu --set full` + source-level stall sites of ONE kernel, summarised with profiles/ncu_summary.py and ncu_source_top.py.
# usage (on the GPU box, via gpurun): scripts/gpu_ncu_kernel.sh <tag> <kernel-regex> <skip> -- <command ...>
#   e.g. scripts/gpu_ncu_kernel.sh r2a apply_sparse_kernel 2 -- python bench.py --workload apply_sparse --nseq 50000 --steps 1 --warmup 1 --no-cpu --no-e2e
TAG=$1; RE=$2; SKIP=$3; shift 4
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c 1 -o $OUT/prof "$@" > $OUT/ncu.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/prof_source.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_raw.csv | cut -c1-150
python profiles/ncu_source_top.py $OUT/prof_source.csv 30 | cut -c1-200
