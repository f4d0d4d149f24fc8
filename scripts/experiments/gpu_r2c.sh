#!/bin/bash
# R2c: the rewritten count_dense_warp_kernel — parity tests that touch it, A/B against the round-1 kernel, ncu --set full.
TAG=${1:-R2c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_full_size.py tests/test_gpu_property.py tests/test_api_rules.py tests/test_gpu_c1.py -m gpu -q --timeout 600 2>&1 | tail -30 > $OUT/pytest.txt; tail -30 $OUT/pytest.txt
SKM_CDW_BYTES=1 timeout 300 python bench.py --workload vectorize --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/vec_bytes.json 2> $OUT/vec_bytes.err
timeout 300 python bench.py --workload vectorize --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/vec_new.json 2> $OUT/vec_new.err
python - <<PY
import json
for f in ("vec_bytes","vec_new"):
    d=json.load(open("$OUT/%s.json"%f)); print(f, d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:count_dense_warp_kernel -s 2 -c 1 -o $OUT/prof_cdw python bench.py --workload vectorize --steps 3 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_cdw.log 2>&1
ncu -i $OUT/prof_cdw.ncu-rep --page raw --csv > $OUT/prof_cdw_raw.csv 2>/dev/null
ncu -i $OUT/prof_cdw.ncu-rep --page source --csv > $OUT/prof_cdw_source.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_cdw_raw.csv | cut -c1-150 | grep -v "tensor" > $OUT/ncu_full_count_dense_warp_summary.txt; cat $OUT/ncu_full_count_dense_warp_summary.txt
python profiles/ncu_source_top.py $OUT/prof_cdw_source.csv 25 | cut -c1-200 > $OUT/count_dense_warp_stalls.txt; head -30 $OUT/count_dense_warp_stalls.txt
