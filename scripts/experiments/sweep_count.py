#!/usr/bin/env python
"""Time skm_count_dense (C2 workload) for several (threads, rows) tile shapes.  GPU box only."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from bench import synth_proteins
from snekmer_b200 import engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
res, off = synth_proteins(n, 2)
batch = E.SequenceBatch.from_packed(res, off)
basis = E.build_basis(batch, "miqs", 3, 0)
out = torch.empty((batch.n, basis.K), dtype=torch.int32, device="cuda")
ref = None
alg = batch.nres + 8 * (batch.n + 1) + 4 * batch.n * basis.K
for threads, rows in [tuple(map(int, x.split("x"))) for x in (sys.argv[2] if len(sys.argv) > 2 else "256x12,256x8,128x4").split(",")]:
    os.environ["SKM_CD_THREADS"], os.environ["SKM_CD_ROWS"] = str(threads), str(rows)
    for _ in range(3):
        E.count_dense(batch, "miqs", 3, basis, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        E.count_dense(batch, "miqs", 3, basis, out=out)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    chk = int(out.sum().item())
    ref = chk if ref is None else ref
    print(json.dumps({"threads": threads, "rows": rows, "ms": round(ms, 4), "GBs": round(alg / ms / 1e6, 1), "ok": chk == ref}), flush=True)
