#!/bin/bash
OUT=gpurun_out/r1i; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x -k "not tensor_cores" 2>&1 | tail -40 > $OUT/pytest.txt; tail -15 $OUT/pytest.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "warp_sort and (5-3 or None-8)" > $OUT/memcheck.txt 2>&1; echo memcheck rc=$?; tail -4 $OUT/memcheck.txt
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "warp_sort and 5-3" > $OUT/racecheck.txt 2>&1; echo racecheck rc=$?; tail -4 $OUT/racecheck.txt
timeout 900 python bench.py --workload sweep --steps 3 --warmup 3 > $OUT/bench_sweep.json 2> $OUT/bench_sweep.err
echo "== sweep rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r1i/bench_sweep.json'))
print(d['value'], d['ms_per_step'])
for p in d['config']['points']: print(p['alphabet'],p['k'],p['path'],p['K'],round(p['ms'],3),round(p['hbm_frac'],4))
print(d.get('cpu_baseline'))
PY
tail -3 $OUT/bench_sweep.err
timeout 600 python bench.py --workload apply_sparse --steps 3 --warmup 3 > $OUT/bench_apply_sparse.json 2> $OUT/bench_apply_sparse.err; cut -c1-400 $OUT/bench_apply_sparse.json
