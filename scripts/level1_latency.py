#!/usr/bin/env python
"""Per-call cost of the level-1 drop-in KmerVec.reduce_vectorize(seq) (one H2D + launch + D2H + sync per sequence) —
the number INTEGRATION.md quotes next to the reference's pure-Python call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snekmer_b200.vectorize import KmerVec

rng = np.random.default_rng(0)
aa = list("ACDEFGHIKLMNPQRSTVWY")
seqs = ["".join(rng.choice(aa, size=350)) for _ in range(300)]
for a, k in ((5, 3), (2, 8)):
    kv = KmerVec(alphabet=a, k=k)
    for s in seqs[:20]:
        kv.reduce_vectorize(s)
    t = time.perf_counter()
    for s in seqs:
        kv.reduce_vectorize(s)
    dt = (time.perf_counter() - t) / len(seqs)
    print("snekmer_b200 reduce_vectorize alphabet", a, "k", k, ": %.1f us per 350-residue sequence" % (dt * 1e6))
