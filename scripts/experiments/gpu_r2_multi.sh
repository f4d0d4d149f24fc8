#!/bin/bash
# Round 2, N-GPU box: NCCL exchange-step test, then the DEFAULT bench line (all workloads, collectives, in-run parity) at N ranks.
# usage: scripts/gpu_r2_multi.sh <tag> <N>
TAG=${1:-multi}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_dist.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_default_n$N.json 2> $OUT/bench_default_n$N.err
echo "== default N=$N rc=$?"; tail -5 $OUT/bench_default_n$N.err
python - <<PY
import json
d=json.load(open("$OUT/bench_default_n$N.json"))
print("vectorize", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("pcie_gbs_per_rank"), "parity", d.get("parity_check"))
for k,w in d["workloads"].items():
    if "error" in w: print(k, "ERROR", w["error"], w["traceback"][-800:]); continue
    print(k, w["value"], w["ms_per_step"], "comm", w.get("comm_ms"), "e2e", w.get("e2e",{}).get("value"), "|", w.get("parity_check"))
    if w.get("annotation_sharded"): print("   sharded", w["annotation_sharded"]["ms_per_step"], w["annotation_sharded"]["comm_ms"])
PY
