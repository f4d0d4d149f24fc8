#!/bin/bash
# on-chip learn (skm_ann_sort): parity tests, learn bench at N=1 for onchip / hybrid, launch list + ncu of rows_accumulate / ann_sort
TAG=${1:-R2q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_round2.py -k "onchip or hybrid or c3_shaped" -x -q > $OUT/pytest_learn.txt 2>&1; tail -15 $OUT/pytest_learn.txt
run() { name=$1; shift; env "$@" timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --workload learn --no-e2e > $OUT/learn_$name.json 2> $OUT/learn_$name.err; python -c "import json;d=json.load(open('$OUT/learn_$name.json'));print('$name', d['ms_per_step'], d['parity_check'][:30], d['local_learn_ms_events_this_rank'], d['untimed_local_learn_ms_events_this_rank']); print({k:v for k,v in d['launches_per_step'].items() if 'skm' in k or 'cub' in k.lower()})" || tail -5 $OUT/learn_$name.err; }
run onchip SKM_LEARN_METHOD=onchip
run onchip_s05 SKM_LEARN_METHOD=onchip SKM_ONCHIP_HEAVY_SCALE=0.5
run onchip_s12 SKM_LEARN_METHOD=onchip SKM_ONCHIP_HEAVY_SCALE=1.2
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_learn.csv python bench.py --gpus 1 --steps 1 --warmup 3 --workload learn --no-e2e > $OUT/ncu_launch.log 2>&1
python profiles/launch_summary.py $OUT/launches_learn.csv > $OUT/launches_learn_summary.txt 2>&1 || true
head -30 $OUT/launches_learn_summary.txt
# ncu --set full of the biggest rows_accumulate launch (the unannotated row: last launch group) and of ann_sort_kernel
bash scripts/gpu_ncu_kernel.sh $TAG/ra rows_accumulate_kernel 0 -- python bench.py --gpus 1 --steps 1 --warmup 0 --workload learn --no-e2e --no-cpu > $OUT/ncu_ra_summary.txt 2>&1
bash scripts/gpu_ncu_kernel.sh $TAG/as ann_sort_kernel 0 -- python bench.py --gpus 1 --steps 1 --warmup 0 --workload learn --no-e2e --no-cpu > $OUT/ncu_as_summary.txt 2>&1
grep -i "duration\|dram__bytes\|lts__t_sect.*red\|atom\|issue_active\|inst_executed.sum \|stall" $OUT/ncu_ra_summary.txt | head -30
