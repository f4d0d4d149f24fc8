#!/bin/bash
# the driver's command at N ranks: default line (all workloads), then the reference arm
TAG=${1:-R2p}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
S=$(date +%s)
if [ "$N" = 1 ]; then
  timeout 1200 python bench.py --gpus 1 > $OUT/bench_default_n$N.json 2> $OUT/bench_default_n$N.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N > $OUT/bench_default_n$N.json 2> $OUT/bench_default_n$N.err
fi
echo "rc=$? wall=$(( $(date +%s) - S )) s"
python - <<P
import json
d=json.loads(open('$OUT/bench_default_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('e2e',{}).get('value'), d.get('parity_check'))
for k,v in d['workloads'].items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('comm_ms'), str(v.get('parity_check'))[:50], (v.get('e2e') or {}).get('value'), v.get('error'))
P
tail -5 $OUT/bench_default_n$N.err
