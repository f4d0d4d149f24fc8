#!/bin/bash
# learn_host test + learn bench (with e2e) for several heavy-row thresholds of the hybrid path
TAG=${1:-R2aa}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_round2.py -k "learn_host or hybrid" -x -q > $OUT/pytest.txt 2>&1; tail -3 $OUT/pytest.txt
for d in 4 8 12 20; do
  SKM_HEAVY_ROW_DIV=$d timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --workload learn --no-cpu > $OUT/learn_div$d.json 2> $OUT/learn_div$d.err
  python -c "import json;d=json.load(open('$OUT/learn_div$d.json'));print('div $d', d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step'])" || tail -5 $OUT/learn_div$d.err
done
