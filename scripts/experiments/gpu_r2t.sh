#!/bin/bash
# branch-free sparse apply + 16-byte peer stores: parity tests, apply_sparse bench at N=1, learn at N=2
TAG=${1:-R2t}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_rules_sparse.py -k "sparse or c4" -x -q > $OUT/pytest_sparse.txt 2>&1; tail -4 $OUT/pytest_sparse.txt
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q > $OUT/pytest_dist.txt 2>&1; tail -3 $OUT/pytest_dist.txt
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --workload apply_sparse --no-e2e --no-cpu > $OUT/apply_sparse_n1.json 2> $OUT/apply_sparse_n1.err
python -c "import json;d=json.load(open('$OUT/apply_sparse_n1.json'));print('apply_sparse', d['ms_per_step'], d['value'], d['parity_check'][:60])" || tail -5 $OUT/apply_sparse_n1.err
if [ "$N" -gt 1 ]; then
env SKM_EXCHANGE=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 --workload learn --no-e2e > $OUT/learn_peer.json 2> $OUT/learn_peer.err
python -c "import json;d=json.load(open('$OUT/learn_peer.json'));print('learn', d['ms_per_step'], d['comm_ms'], d['comm_phases_ms_this_rank'], d['parity_check'][:40])" || tail -5 $OUT/learn_peer.err
fi
