#!/bin/bash
# light list of the hybrid learn written straight to its final places (skm_learn_sparse_group_place): parity + bench A/B
TAG=${1:-R2ah}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_rules_sparse.py tests/test_gpu_kernels.py -k "learn or hybrid or c3_shaped or rules_sparse" -x -q > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
for m in 1 0; do
  SKM_LEARN_PLACE=$m timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --workload learn --no-cpu --no-e2e > $OUT/learn_place$m.json 2> $OUT/learn_place$m.err
  python -c "import json;d=json.load(open('$OUT/learn_place$m.json'));print('place $m', d['ms_per_step'], d['parity_check'][:40])" || tail -5 $OUT/learn_place$m.err
done
