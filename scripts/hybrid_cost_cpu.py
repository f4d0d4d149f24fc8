"""CPU-only (numpy) cost model of a hybrid tensor-core / SpMM sparse apply on the C4 shape: learns the 400 k-protein matrix
with numpy, takes 20 k queries, ranks the k-mer columns by their multiply-accumulates f_j * c_j (f_j = queries holding the
k-mer, c_j = annotations holding it) and prints, per number H of columns moved to a dense int8 GEMM, the share of the
multiply-accumulates they hold and the resulting time estimate (GEMM at the measured int8 rate + the SpMM tail at today's
80 ms per 200 k queries).  Also prints the column-length percentiles and the lane utilisation of the SpMM's segments.
Takes ~5 minutes and ~10 GB of host memory.  usage: python scripts/hybrid_cost_cpu.py"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from oracle import skm_oracle as O
SYN6 = bench.SYN6_ORACLE if hasattr(bench, "SYN6_ORACLE") else None
t0=time.time()
n_ann, k = 50000, 8
tr_res, tr_off = bench.synth_proteins(400000, 79)
tr_ann = bench.zipf_annotations(400000, n_ann, 0.0, 80)
# LUT for syn6
groups = bench.SYN6
syms = sorted(set(groups.values()))
lut = np.full(256, 255, np.uint8)
for g, s in groups.items():
    for ch in g: lut[ord(ch)] = syms.index(s)
nsym = len(syms); S = nsym ** k
sym = lut[tr_res]
# rolling codes
n = len(sym)
code = np.zeros(n, np.int64); bad = np.zeros(n, bool)
for j in range(k):
    sh = np.roll(sym, -j).astype(np.int64)
    code = code * nsym + np.where(sh == 255, 0, sh)
    bad |= (sh == 255)
# windows crossing sequence ends invalid
seq_of = np.repeat(np.arange(len(tr_off) - 1), np.diff(tr_off))
endpos = tr_off[1:][seq_of]
pos = np.arange(n)
valid = (~bad) & (pos + k <= endpos)
key = tr_ann[seq_of[valid]].astype(np.int64) * S + code[valid]
print("windows", valid.sum(), time.time()-t0)
uk = np.unique(key)
print("nnz(M)", len(uk), time.time()-t0)
col = uk % S
lens = np.bincount(col, minlength=S)
# queries
res, off = bench.synth_proteins(20000, 5)
symq = lut[res]; nq=len(symq)
codeq = np.zeros(nq, np.int64); badq = np.zeros(nq, bool)
for j in range(k):
    sh = np.roll(symq, -j).astype(np.int64)
    codeq = codeq * nsym + np.where(sh == 255, 0, sh); badq |= (sh == 255)
seqq = np.repeat(np.arange(len(off) - 1), np.diff(off))
validq = (~badq) & (np.arange(nq) + k <= off[1:][seqq])
qk = np.unique(seqq[validq].astype(np.int64) * S + codeq[validq])
ql = lens[qk % S]
Q = len(off) - 1
print("entries/query", len(qk)/Q, "MACs/query", ql.sum()/Q)
for pct in (10, 25, 50, 75, 90, 99): print("pct", pct, np.percentile(ql, pct))
for seg in (512, 256, 128, 64, 32):
    slots = np.ceil(ql / seg).sum() * seg
    print("seg", seg, "lane-slots/query", slots / Q, "utilisation", ql.sum() / slots)
qc = qk % S
f = np.bincount(qc, minlength=len(lens)).astype(np.float64) * (200000 / Q)   # per 200k queries
mac = f * lens
tot = mac.sum()
print("total MACs per 200k queries", tot / 1e9, "G")
order = np.argsort(-mac)
cs = np.cumsum(mac[order])
for H in (256, 512, 1024, 2048, 4096, 8192, 16384):
    share = cs[H - 1] / tot
    dense_ms = H * 200000 * 50000 / 1.6e15 * 1e3
    print(H, f"columns: share of MACs {share:.3f}, dense GEMM {dense_ms:.1f} ms, SpMM tail {(1 - share) * 80:.1f} ms")
