#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench, --set full captures of the dominant kernels.
# usage: scripts/gpu_profiles.sh <tag>
TAG=${1:-prof}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_vectorize.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_launch_vectorize.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_apply.csv python bench.py --workload apply --nseq 200000 --steps 2 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_launch_apply.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"count_dense_kernel|basis_kernel" -s 6 -c 2 -o $OUT/vectorize python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_full_vectorize.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_tc_kernel -s 1 -c 1 -o $OUT/apply_tc python scripts/apply_micro.py --mmax 200 --reps 1 > $OUT/ncu_full_apply_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_sparse_kernel -s 1 -c 1 -o $OUT/apply_sparse python bench.py --workload apply_sparse --nseq 20000 --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_full_apply_sparse.log 2>&1
ls -la $OUT; tail -2 $OUT/*.log
