#!/bin/bash
# N-rank learn with different NCCL p2p channel settings (all_to_all of the COO runs: 5.45 ms of the 8.7 ms exchange at N=2)
TAG=${1:-R2m}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 --workload learn --no-e2e > $OUT/learn_$name.json 2> $OUT/learn_$name.err; python -c "import json;d=json.load(open('$OUT/learn_$name.json'));print('$name', d['ms_per_step'], d['comm_ms'], d['comm_phases_ms_this_rank'])" || tail -3 $OUT/learn_$name.err; }
run default SKM_X=0
run p2p32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run p2p64 NCCL_MIN_P2P_NCHANNELS=64 NCCL_MAX_P2P_NCHANNELS=64 NCCL_MAX_NCHANNELS=64
run ctas NCCL_MIN_CTAS=32 NCCL_MIN_P2P_NCHANNELS=32
