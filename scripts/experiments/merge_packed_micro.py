"""Time skm_coo_merge_runs_packed on W sorted runs of packed words shaped like the N = 8 learn fan-in (120 M words in 8 runs,
~22 % duplicate keys across runs).  usage: python scripts/experiments/merge_packed_micro.py [W] [words_per_run]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from snekmer_b200 import engine as E

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
per = int(sys.argv[2]) if len(sys.argv) > 2 else 15_000_000
bits = 29
g = torch.Generator(device="cuda").manual_seed(1)
runs = []
for r in range(W):
    k = torch.randint(0, int(per * W / 1.3), (int(per * 1.12),), device="cuda", generator=g, dtype=torch.int64)
    k = torch.unique(k)[:per]                       # sorted unique keys; overlap between runs ~ like the fan-in
    runs.append((k << bits) | torch.randint(1, 50, (k.numel(),), device="cuda", generator=g, dtype=torch.int64))
sizes = [int(x.numel()) for x in runs]
buf = torch.cat(runs)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(4):
    ev[0].record()
    ok, ov, dn = E.coo_merge_runs_packed(buf.data_ptr(), sizes, bits, buf.device)
    ev[1].record(); ev[1].synchronize()
    m = int(dn.item())
    print(f"W={W} words={sum(sizes)} -> entries={m}  {ev[0].elapsed_time(ev[1]):.3f} ms")
# check against torch
keys = buf >> bits
uk, inv = torch.unique(keys, return_inverse=True)
uv = torch.zeros_like(uk).index_add_(0, inv, buf & ((1 << bits) - 1))
assert m == uk.numel() and torch.equal(ok[:m], uk) and torch.equal(ov[:m], uv)
print("ok")
