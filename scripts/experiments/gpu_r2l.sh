#!/bin/bash
# N-rank run of learn and apply_sparse alone (exchange phase breakdown, MAC-balanced annotation ranges)
TAG=${1:-R2l}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for w in learn apply_sparse; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 --workload $w --no-e2e > $OUT/n${N}_$w.json 2> $OUT/n${N}_$w.err
  echo "== $w N=$N rc=$?"; tail -2 $OUT/n${N}_$w.err
done
python - <<PY
import json
d=json.load(open("$OUT/n${N}_learn.json")); print("learn", d["ms_per_step"], d["comm_ms"], d["comm_phases_ms_this_rank"], d["parity_check"][:40])
d=json.load(open("$OUT/n${N}_apply_sparse.json")); print("apply_sparse", d["ms_per_step"], d["roofline"]["kernel_ms"], json.dumps(d["annotation_sharded"])[:700], d["parity_check"][:30])
PY
