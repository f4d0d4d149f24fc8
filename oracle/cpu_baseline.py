"""CPU baseline = the oracle port run on the box's host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/skm_oracle.py).  Used by
bench.py's ``cpu_baseline`` leg and by ``bench.py --impl reference``; never by
the product.  Mirrors how the reference parallelises (one Snakemake job per
input shard, cli.py:134-150): the sample is cut into contiguous shards, one
worker process per shard, two phases like the vectorize rule
(kmerize.smk:89-104 basis pass, :112-120 count pass)."""
from __future__ import annotations

import os
import time
from multiprocessing import get_context

import numpy as np

from . import skm_oracle as O

_G = {}


def _init(residues, offsets, alphabet, k):
    _G["res"], _G["off"], _G["a"], _G["k"] = residues, offsets, alphabet, k
    _G["lut"], _G["syms"] = O.build_lut(alphabet)


def _shard(lo, hi):
    off = _G["off"][lo:hi + 1]
    return _G["res"][off[0]:off[-1]], off - off[0], int(off[0])


def _phase1(args):
    lo, hi = args
    res, off, base = _shard(lo, hi)
    si, pos, code, valid = O.window_codes(res, off, _G["lut"], len(_G["syms"]), _G["k"])
    c = code[valid]
    g = (off[si] + pos)[valid] + base
    uniq, first, cnt = np.unique(c, return_index=True, return_counts=True)
    return uniq, g[first], cnt


def _phase2(args):
    lo, hi, basis = args
    res, off, _ = _shard(lo, hi)
    si, pos, code, valid = O.window_codes(res, off, _G["lut"], len(_G["syms"]), _G["k"])
    C = O.count_matrix(si, code, valid, hi - lo, basis)
    return int(C.sum())       # the matrix stays in the worker (the reference writes it to a file)


def vectorize_sample(residues, offsets, alphabet, k, workers=None, min_filter=0):
    """Time the two-pass vectorize on (residues, offsets).  Returns dict(seconds, nseq, cores, K)."""
    n = len(offsets) - 1
    workers = max(1, min(workers or os.cpu_count() or 1, 64, n))
    cuts = np.linspace(0, n, workers + 1).astype(np.int64)
    shards = [(int(cuts[i]), int(cuts[i + 1])) for i in range(workers) if cuts[i + 1] > cuts[i]]
    ctx = get_context("fork")
    _init(residues, offsets, alphabet, k)
    with ctx.Pool(len(shards)) as pool:
        pool.map(_phase1, shards[:1])                      # warm the workers (imports, page faults)
        t0 = time.perf_counter()
        parts = pool.map(_phase1, shards)
        codes = np.concatenate([p[0] for p in parts])
        first = np.concatenate([p[1] for p in parts])
        cnt = np.concatenate([p[2] for p in parts])
        order = np.argsort(codes, kind="stable")
        codes, first, cnt = codes[order], first[order], cnt[order]
        starts = np.flatnonzero(np.r_[True, codes[1:] != codes[:-1]])
        ucodes = codes[starts]
        ufirst = np.minimum.reduceat(first, starts)
        ucnt = np.add.reduceat(cnt, starts)
        keep = ucnt > min_filter
        basis = ucodes[keep][np.argsort(ufirst[keep], kind="stable")]
        tot = pool.map(_phase2, [(lo, hi, basis) for lo, hi in shards])
        dt = time.perf_counter() - t0
    return {"seconds": dt, "nseq": n, "cores": len(shards), "K": int(len(basis)), "checksum": int(sum(tot))}
