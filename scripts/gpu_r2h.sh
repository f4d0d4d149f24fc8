#!/bin/bash
OUT=gpurun_out/r2h; mkdir -p $OUT
cat > /tmp/cs.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from snekmer_b200 import engine as E
res, off = bench.synth_proteins(200000, 5)
b = E.SequenceBatch.from_packed(res, off)
for _ in range(3): E.count_csr(b, "miqs", 6)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_sort_kernel" -s 2 -c 1 -o $OUT/prof_cs python /tmp/cs.py > $OUT/ncu_cs.log 2>&1
ncu -i $OUT/prof_cs.ncu-rep --page raw --csv > $OUT/prof_cs_raw.csv 2>/dev/null
ncu -i $OUT/prof_cs.ncu-rep --page source --csv > $OUT/prof_cs_source.csv 2>/dev/null
python profiles/ncu_summary.py $OUT/prof_cs_raw.csv | cut -c1-150
python profiles/ncu_source_top.py $OUT/prof_cs_source.csv 26 | cut -c1-200
