"""Micro-benchmark of skm_apply_tc (tcgen05 scoring GEMM + fused top-2) on random operands.
usage: python scripts/apply_micro.py [--nq N] [--ann A] [--K K] [--mmax M] [--reps R]
Prints kernel ms (CUDA events, split_q + apply_tc) and the int8 TOP/s actually issued
(sum over annotation tiles of their digit planes)."""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snekmer_b200 import engine as E

ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, default=148 * 2 * 128)
ap.add_argument("--ann", type=int, default=50000)
ap.add_argument("--K", type=int, default=1000)
ap.add_argument("--mmax", type=int, default=200)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
Q = torch.randint(0, 4, (a.nq, a.K), generator=g, device=dev, dtype=torch.int32)
M = torch.randint(0, a.mmax + 1, (a.ann, a.K), generator=g, device=dev, dtype=torch.int64)
prep = E.prepare_annotations(M)
tiles = prep.tiles()
n_tiles = len(tiles)
qn2 = E.row_norm2(Q)
for _ in range(2):
    r = E.apply_tc(Q, prep, qn2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    r = E.apply_tc(Q, prep, qn2)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
Kp = (a.K + 127) // 128 * 128
ops = 2.0 * a.nq * prep.issued_macs_per_query()
print(f"nq={a.nq} ann={a.ann} K={a.K} mmax={a.mmax} planes(max)={prep.n_planes} tiles={n_tiles} mean planes/tile={tiles[:, 2].mean():.2f} "
      f"ms={ms:.3f} TOP/s={ops / ms / 1e9:.1f}")
