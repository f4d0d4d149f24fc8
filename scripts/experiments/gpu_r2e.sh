#!/bin/bash
# R2e: zero-fill variants of the count kernel (each in its own process: a wrong instruction form kills the context) + hybrid learn.
TAG=${1:-R2e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for z in off smem smem_nocluster l2; do
  SKM_CDW_ZFILL=$z timeout 300 python bench.py --workload vectorize --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/vec_z_$z.json 2> $OUT/vec_z_$z.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/vec_z_$z.json")); print("$z", d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac'])
except Exception as e: print("$z FAILED", open("$OUT/vec_z_$z.err").read()[-300:])
PY
done
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py tests/test_gpu_rules_sparse.py tests/test_gpu_full_size.py -m gpu -q --timeout 600 2>&1 | tail -30 > $OUT/pytest.txt; tail -12 $OUT/pytest.txt
timeout 600 python bench.py --workload learn --steps 5 --warmup 3 --no-cpu > $OUT/learn.json 2> $OUT/learn.err
python -c "import json;d=json.load(open('$OUT/learn.json'));print('learn', d['ms_per_step'], d['parity_check'], d['launches_per_step'])" || tail -5 $OUT/learn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_learn.csv python bench.py --workload learn --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_learn.log 2>&1
python profiles/launch_summary.py $OUT/launches_learn.csv > $OUT/launches_learn_summary.txt 2>&1; head -30 $OUT/launches_learn_summary.txt
