#!/bin/bash
# why is the 1-ms vectorize step 0.35 ms slower at 8 ranks than at 1?  vectorize-only runs with one thing switched off each
TAG=${1:-R2ac}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --steps 20 --warmup 5 --workload vectorize --no-e2e --no-cpu > $OUT/vec_$name.json 2> $OUT/vec_$name.err; python -c "import json;d=json.load(open('$OUT/vec_$name.json'));print('$name', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['cuda_graph'][:12], d['clocks'])" || tail -3 $OUT/vec_$name.err; }
run default A=1
run noclocks SKM_NO_CLOCKS=1
run nograph SKM_NO_GRAPH=1
run nograph_noclocks SKM_NO_GRAPH=1 SKM_NO_CLOCKS=1
run nobind SKM_NO_CPU_BIND=1
N_STEPS=100; env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --steps 100 --warmup 5 --workload vectorize --no-e2e --no-cpu > $OUT/vec_steps100.json 2> $OUT/vec_steps100.err; python -c "import json;d=json.load(open('$OUT/vec_steps100.json'));print('steps100', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))"
