#!/bin/bash
OUT=gpurun_out/r1n; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_confidence.py tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "not tensor_cores" 2>&1 | tail -15 > $OUT/pytest.txt; tail -5 $OUT/pytest.txt
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from snekmer_b200 import engine as E
res, off = bench.synth_proteins(200000, 5)
b = E.SequenceBatch.from_packed(res, off)
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for a, k in [("miqs", 6), (None, 8), (None, 14)]:
    wb = E.build_basis_wide(b, a, k, 0)
    print(a, k, "K", wb.K, "csr64 no basis %.2f ms" % t(lambda: E.count_csr_wide(b, a, k)),
          "csr64 + lookup %.2f ms" % t(lambda: E.count_csr_wide(b, a, k, wb)))
PY
timeout 600 python bench.py --workload sweep --steps 3 --warmup 3 --no-cpu > $OUT/bench_sweep.json 2> $OUT/bench_sweep.err; cut -c1-200 $OUT/bench_sweep.json
