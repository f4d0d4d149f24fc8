#!/bin/bash
# N-GPU box: NCCL exchange-step test, then the bench at N ranks (vectorize = headline; learn = the exchange workload).
# usage: scripts/gpu_multi.sh <tag> <N>
TAG=${1:-multi}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_dist.txt
for w in vectorize learn apply; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 --workload $w > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
  echo "== $w N=$N rc=$?"; cut -c1-700 $OUT/bench_${w}_n$N.json; tail -3 $OUT/bench_${w}_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
echo "== reference arm rc=$?"; cut -c1-500 $OUT/bench_ref_n$N.json; tail -3 $OUT/bench_ref_n$N.err
