#!/bin/bash
OUT=gpurun_out/r1x; mkdir -p $OUT
cat > /tmp/bb.py <<'PY'
import sys, os, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from snekmer_b200 import engine as E
res, off = bench.synth_proteins(1000000, 2)
b = E.SequenceBatch.from_packed(res, off)
def t(f, n=10):
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
count, first = E.basis_tables(1000, b.device)
for mode in ("0", "1"):
    os.environ["SKM_BASIS_TILE"] = mode
    print("tile" if mode == "1" else "warp", "basis_accumulate %.3f ms" % t(lambda: E.basis_accumulate(b, "miqs", 3, count, first, 0)),
          " counts only %.3f ms" % t(lambda: E.basis_accumulate(b, "miqs", 3, count, None, 0)))
os.environ.pop("SKM_BASIS_TILE")
PY
python /tmp/bb.py
timeout 600 ncu --set full --clock-control none -k regex:"basis_warp_kernel" -s 1 -c 1 -o $OUT/prof_bw python /tmp/bb.py > $OUT/ncu_bw.log 2>&1
ncu -i $OUT/prof_bw.ncu-rep --page raw --csv > $OUT/prof_bw_raw.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_bw_raw.csv | cut -c1-150
