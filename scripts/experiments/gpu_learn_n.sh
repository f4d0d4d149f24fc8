#!/bin/bash
# usage: scripts/gpu_learn_n.sh <tag> <N>: NCCL exchange test + learn bench at N ranks
TAG=${1:-ln}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 600 2>&1 | tail -8 | tee $OUT/pytest_dist.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 --workload learn --no-e2e > $OUT/bench_learn_n$N.json 2> $OUT/bench_learn_n$N.err
echo "== learn N=$N rc=$?"; cut -c1-400 $OUT/bench_learn_n$N.json; grep -v "OMP_NUM\|^\*\*\*" $OUT/bench_learn_n$N.err | tail -5
