#!/bin/bash
# tensor-core scoring: parity tests, micro-benchmark sweep, optional ncu capture.  usage: scripts/gpu_apply_prof.sh <tag> [ncu]
TAG=${1:-run}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -k "tensor_cores or apply" 2>&1 | tail -15 | tee $OUT/pytest_tc.txt
for cfg in "--mmax 200" "--mmax 70000" "--mmax 200 --ann 4096" "--mmax 200 --K 256" "--mmax 200 --K 4096 --nq 18944"; do
  timeout 120 python scripts/apply_micro.py $cfg 2>&1 | tail -1
done | tee $OUT/apply_micro.txt
if [[ "$2" == ncu ]]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_tc_kernel -s 1 -c 1 -o $OUT/apply_tc python scripts/apply_micro.py --mmax 200 --reps 1 > $OUT/ncu_apply.log 2>&1
  tail -3 $OUT/ncu_apply.log
fi
