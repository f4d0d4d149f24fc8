#!/bin/bash
# CUDA-graph front of the vectorize step: test + vectorize bench with / without the graph
TAG=${1:-R2v}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_round2.py -k "order_only" -x -q > $OUT/pytest.txt 2>&1; tail -5 $OUT/pytest.txt
for g in graph nograph; do
  if [ $g = nograph ]; then export SKM_NO_GRAPH=1; else unset SKM_NO_GRAPH; fi
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workload vectorize --no-cpu --no-e2e > $OUT/vec_$g.json 2> $OUT/vec_$g.err
  python -c "import json;d=json.load(open('$OUT/vec_$g.json'));print('$g', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['whole_step']['frac'], d['cuda_graph'][:20], d['launches_per_step'])" || tail -5 $OUT/vec_$g.err
done
