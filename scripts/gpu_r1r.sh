#!/bin/bash
OUT=gpurun_out/r1r; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "csr or sparse or wide or dispatch" 2>&1 | tail -15 > $OUT/pytest.txt; tail -5 $OUT/pytest.txt
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
print("csr32 miqs6 %.2f ms" % t(lambda: E.count_csr(b, "miqs", 6)), " hydro14 %.2f ms" % t(lambda: E.count_csr(b, "hydro", 14)))
print("csr64 None14 no basis %.2f ms" % t(lambda: E.count_csr_wide(b, None, 14)))
PY
