#!/bin/bash
# R2b: A/B of the count kernel's scan (bytes vs words), ncu --set full of the new kernel and of apply_tc_kernel with
# the tensor-pipe metrics, launch list of the default bench line.
TAG=${1:-R2b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "kmerlist_restriction or dense or transports or golden" 2>&1 | tail -5
SKM_CDW_BYTES=1 timeout 300 python bench.py --workload vectorize --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/vec_bytes.json 2> $OUT/vec_bytes.err
echo "== bytes"; python -c "import json;d=json.load(open('$OUT/vec_bytes.json'));print(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
timeout 300 python bench.py --workload vectorize --steps 20 --warmup 5 --no-cpu > $OUT/vec_words.json 2> $OUT/vec_words.err
echo "== words"; python -c "import json;d=json.load(open('$OUT/vec_words.json'));print(d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['ms_per_step'],[ (k[:6],v['value']) for k,v in d['e2e']['transports'].items()])"
timeout 300 python bench.py --workload apply --steps 3 --warmup 3 --no-cpu > $OUT/apply.json 2> $OUT/apply.err
echo "== apply"; python -c "import json;d=json.load(open('$OUT/apply.json'));print(d['ms_per_step'],d['e2e']['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_dense_warp_kernel -s 2 -c 1 -o $OUT/prof_cdw python bench.py --workload vectorize --steps 3 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_cdw.log 2>&1
ncu -i $OUT/prof_cdw.ncu-rep --page raw --csv > $OUT/prof_cdw_raw.csv 2>/dev/null
ncu -i $OUT/prof_cdw.ncu-rep --page source --csv > $OUT/prof_cdw_source.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_cdw_raw.csv | cut -c1-150 > $OUT/ncu_full_count_dense_warp_words_summary.txt; cat $OUT/ncu_full_count_dense_warp_words_summary.txt
python profiles/ncu_source_top.py $OUT/prof_cdw_source.csv 25 | cut -c1-200 > $OUT/count_dense_warp_words_stalls.txt
timeout 900 ncu --set full --clock-control none -k regex:apply_tc_kernel -s 1 -c 1 -o $OUT/prof_tc python bench.py --workload apply --nseq-apply 200000 --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_tc.log 2>&1
ncu -i $OUT/prof_tc.ncu-rep --page raw --csv > $OUT/prof_tc_raw.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_tc_raw.csv | cut -c1-170 > $OUT/ncu_full_apply_tc_summary.txt; cat $OUT/ncu_full_apply_tc_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_default.csv python bench.py --steps 2 --warmup 1 --sub-steps 1 --no-cpu --no-e2e > $OUT/ncu_launch.log 2>&1
python profiles/launch_summary.py $OUT/launches_default.csv > $OUT/launches_default_summary.txt 2>&1; head -40 $OUT/launches_default_summary.txt
rm -f $OUT/prof_tc.ncu-rep
