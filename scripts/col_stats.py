"""C4 shape: how the multiply-accumulates of sparse apply distribute over column lengths (decides whether dense handling
of the heaviest k-mer columns pays).  Prints, per length threshold, the number of columns, their share of the MACs and
the mean number of such columns a query touches."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from snekmer_b200 import alphabet as A
from snekmer_b200 import engine as E

A.register_alphabet("syn6", bench.SYN6)
n_ann, S, k = 50000, 6 ** 8, 8
tr_res, tr_off = bench.synth_proteins(400000, 79)
tr_ann = bench.zipf_annotations(400000, n_ann, 0.0, 80)
tb = E.SequenceBatch.from_packed(tr_res, tr_off)
keys, vals = E.learn_sparse(tb, "syn6", k, torch.from_numpy(tr_ann), n_ann)
csc = E.csc_build(keys, vals, S, n_ann, 0)
res, off = bench.synth_proteins(50000, 5)
qb = E.SequenceBatch.from_packed(res, off)
rowptr, cols, cvals = E.count_csr(qb, "syn6", k, None)
lens = (csc.colptr[1:] - csc.colptr[:-1])
ql = lens[cols.long()]
total = int(ql.sum().item())
nq = qb.n
print("nnz(M)", int(lens.sum().item()), "max entry", csc.max_m, "queries", nq, "MACs/query", total / nq, "entries/query", cols.numel() / nq)
for T in (500, 1000, 2000, 5000, 10000, 20000, 40000):
    heavy = lens > T
    nh = int(heavy.sum().item())
    share = float(ql[ql > T].sum().item()) / total
    per_q = float((ql > T).sum().item()) / nq
    mx = int(csc.mvals[:0].numel())
    print(f"len > {T:6d}: {nh:7d} columns, {100 * share:5.1f} % of MACs, {per_q:6.1f} such entries per query")
# largest value inside heavy columns (uint16 eligibility)
heavy = (lens > 5000).nonzero().reshape(-1)
if heavy.numel():
    seg_max = 0
    for c in heavy[:200].tolist():
        a, b = int(csc.colptr[c].item()), int(csc.colptr[c + 1].item())
        seg_max = max(seg_max, int(csc.mvals[a:b].max().item()))
    print("largest M entry in 200 heavy columns:", seg_max)
