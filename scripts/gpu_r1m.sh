#!/bin/bash
OUT=gpurun_out/r1m; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"count_dense_warp_kernel" -s 3 -c 1 -o $OUT/prof_cdw python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_full_cdw.log 2>&1
ncu -i $OUT/prof_cdw.ncu-rep --page raw --csv > $OUT/prof_cdw_raw.csv 2>/dev/null
ncu -i $OUT/prof_cdw.ncu-rep --page source --csv > $OUT/prof_cdw_source.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_cdw_raw.csv | cut -c1-160
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
    print(a, k, "K", wb.K, "basis_wide %.2f ms" % t(lambda: E.build_basis_wide(b, a, k, 0)),
          "csr64 no basis %.2f ms" % t(lambda: E.count_csr_wide(b, a, k)),
          "csr64 + lookup %.2f ms" % t(lambda: E.count_csr_wide(b, a, k, wb)),
          "local table %.2f ms" % t(lambda: E.basis_table_local(b, a, k, 0)))
    if a == "miqs":
        print("   csr32 %.2f ms" % t(lambda: E.count_csr(b, a, k)))
PY
