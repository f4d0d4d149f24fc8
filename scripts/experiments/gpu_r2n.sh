#!/bin/bash
# peer-memory exchange of sparse learn at N ranks: dist parity test, learn bench peer vs NCCL baseline
TAG=${1:-R2n}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q > $OUT/pytest_dist.txt 2>&1; tail -5 $OUT/pytest_dist.txt
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 --workload learn --no-e2e > $OUT/learn_$name.json 2> $OUT/learn_$name.err; python -c "import json;d=json.load(open('$OUT/learn_$name.json'));print('$name', d['ms_per_step'], d['comm_ms'], d['comm_phases_ms_this_rank'], d['parity_check'][:40], d['local_learn_ms_events_this_rank'], d['step_wall_ms_this_rank'])" || tail -5 $OUT/learn_$name.err; }
run peer SKM_EXCHANGE=peer
run nccl SKM_EXCHANGE=nccl
run ncclwarm SKM_EXCHANGE=nccl SKM_PEER_WARM=1
