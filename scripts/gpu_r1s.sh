#!/bin/bash
OUT=gpurun_out/r1s; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "not tensor_cores" 2>&1 | tail -8 > $OUT/pytest_main.txt; tail -3 $OUT/pytest_main.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "warp_sort and (5-3 or 2-8)" > $OUT/memcheck.txt 2>&1; echo memcheck rc=$?; tail -3 $OUT/memcheck.txt
for w in sweep apply_sparse; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "== $w rc=$?"; cut -c1-260 $OUT/bench_$w.json
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1s/bench_sweep.json'))
for p in d['config']['points']: print(p['alphabet'],p['k'],p['path'],p['K'],round(p['ms'],3),round(p['hbm_frac'],4))
PY
