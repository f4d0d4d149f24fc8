#!/bin/bash
# learn: how many dense rows one rows_accumulate launch should cover (SKM_ROWS_L2_MB)
TAG=${1:-R2al}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for mb in ${MBS:-14 27 40 67 100}; do
  SKM_ROWS_L2_MB=$mb timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --workload learn --no-cpu --no-e2e > $OUT/learn_l2_$mb.json 2> $OUT/learn_l2_$mb.err
  python -c "import json;d=json.load(open('$OUT/learn_l2_$mb.json'));print('L2 MB $mb', round(d['ms_per_step'],3))" || tail -3 $OUT/learn_l2_$mb.err
done
