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


def _phase2_sparse(args):
    lo, hi, basis_sorted = args
    res, off, _ = _shard(lo, hi)
    si, pos, code, valid = O.window_codes(res, off, _G["lut"], len(_G["syms"]), _G["k"])
    j = np.searchsorted(basis_sorted, code)
    j[j >= len(basis_sorted)] = max(len(basis_sorted) - 1, 0)
    keep = valid & (basis_sorted[j] == code) if len(basis_sorted) else np.zeros_like(valid)
    rowptr, codes, cnt = O.count_csr(si, code, keep, hi - lo)
    return int(cnt.sum())


def vectorize_sparse_sample(residues, offsets, alphabet, k, workers=None, min_filter=0):
    """vectorize_sample for bases too large for dense rows: phase 2 produces per-sequence (code, count) lists
    (the alphabet / k sweep).  Returns dict(seconds, nseq, cores, K)."""
    n = len(offsets) - 1
    workers = max(1, min(workers or os.cpu_count() or 1, 64, n))
    cuts = np.linspace(0, n, workers + 1).astype(np.int64)
    shards = [(int(cuts[i]), int(cuts[i + 1])) for i in range(workers) if cuts[i + 1] > cuts[i]]
    ctx = get_context("fork")
    _init(residues, offsets, alphabet, k)
    with ctx.Pool(len(shards)) as pool:
        pool.map(_phase1, shards[:1])
        t0 = time.perf_counter()
        parts = pool.map(_phase1, shards)
        codes = np.concatenate([p[0] for p in parts])
        first = np.concatenate([p[1] for p in parts])
        cnt = np.concatenate([p[2] for p in parts])
        order = np.argsort(codes, kind="stable")
        codes, first, cnt = codes[order], first[order], cnt[order]
        starts = np.flatnonzero(np.r_[True, codes[1:] != codes[:-1]]) if len(codes) else np.zeros(0, np.int64)
        ucodes = codes[starts]
        ucnt = np.add.reduceat(cnt, starts) if len(starts) else cnt[:0]
        basis_sorted = ucodes[ucnt > min_filter]
        tot = pool.map(_phase2_sparse, [(lo, hi, basis_sorted) for lo, hi in shards])
        dt = time.perf_counter() - t0
    return {"seconds": dt, "nseq": n, "cores": len(shards), "K": int(len(basis_sorted)), "checksum": int(sum(tot))}


# ---------------------------------------------------------------------------
# learn / apply samples (same fan-out: one process per shard)
# ---------------------------------------------------------------------------
def _learn_shard(args):
    lo, hi, ann, S = args
    res, off, _ = _shard(lo, hi)
    si, pos, code, valid = O.window_codes(res, off, _G["lut"], len(_G["syms"]), _G["k"])
    a = ann[si]
    keep = valid & (a >= 0)
    keys = a[keep].astype(np.uint64) * np.uint64(S) + code[keep]
    uk, cnt = np.unique(keys, return_counts=True)
    tot = np.bincount(code[valid].astype(np.int64), minlength=S) if S <= (1 << 24) else None
    return uk, cnt.astype(np.int64), tot


def learn_sample(residues, offsets, ann_id, alphabet, k, extra=None, workers=None):
    """Sparse learn (learn.smk:306-326,359-408 + merge :467-494) on (residues, offsets): per-shard
    (annotation, k-mer) counts with numpy, then one merge.  Returns dict(seconds, nseq, cores, nnz)."""
    n = len(offsets) - 1
    workers = max(1, min(workers or os.cpu_count() or 1, 64, n))
    cuts = np.linspace(0, n, workers + 1).astype(np.int64)
    shards = [(int(cuts[i]), int(cuts[i + 1])) for i in range(workers) if cuts[i + 1] > cuts[i]]
    _G["res"], _G["off"], _G["a"], _G["k"] = residues, offsets, alphabet, k
    _G["lut"], _G["syms"] = O.build_lut(alphabet, extra)
    S = len(_G["syms"]) ** k
    ctx = get_context("fork")
    with ctx.Pool(len(shards)) as pool:
        pool.map(_learn_shard, [(lo, min(hi, lo + 8), ann_id[lo:min(hi, lo + 8)], S) for lo, hi in shards[:1]])
        t0 = time.perf_counter()
        parts = pool.map(_learn_shard, [(lo, hi, ann_id[lo:hi], S) for lo, hi in shards])
        keys = np.concatenate([p[0] for p in parts])
        cnt = np.concatenate([p[1] for p in parts])
        order = np.argsort(keys, kind="stable")
        keys, cnt = keys[order], cnt[order]
        starts = np.flatnonzero(np.r_[True, keys[1:] != keys[:-1]]) if keys.size else np.zeros(0, np.int64)
        merged = np.add.reduceat(cnt, starts) if keys.size else cnt
        dt = time.perf_counter() - t0
    return {"seconds": dt, "nseq": n, "cores": len(shards), "nnz": int(len(starts)), "checksum": int(merged.sum())}


def _apply_shard(args):
    lo, hi, basis, Mn = args
    res, off, _ = _shard(lo, hi)
    si, pos, code, valid = O.window_codes(res, off, _G["lut"], len(_G["syms"]), _G["k"])
    C = O.count_matrix(si, code, valid, hi - lo, basis).astype(np.float64)
    qn = np.sqrt((C * C).sum(axis=1))
    qn[qn == 0] = 1.0
    S = (C / qn[:, None]) @ Mn.T                         # sklearn cosine_similarity(M, Q).T
    top = np.argpartition(-S, 1, axis=1)[:, :2] if S.shape[1] > 1 else np.zeros((hi - lo, 1), np.int64)
    return int(top[:, 0].sum())


def apply_dense_sample(residues, offsets, alphabet, k, basis, M, workers=None):
    """Dense apply (apply.smk:188-335): counts over the basis, float64 cosine against M, top-2."""
    n = len(offsets) - 1
    workers = max(1, min(workers or os.cpu_count() or 1, 64, n))
    cuts = np.linspace(0, n, workers + 1).astype(np.int64)
    shards = [(int(cuts[i]), int(cuts[i + 1])) for i in range(workers) if cuts[i + 1] > cuts[i]]
    _init(residues, offsets, alphabet, k)
    Mf = np.asarray(M, dtype=np.float64)
    mn = np.sqrt((Mf * Mf).sum(axis=1))
    mn[mn == 0] = 1.0
    Mn = Mf / mn[:, None]
    ctx = get_context("fork")
    with ctx.Pool(len(shards)) as pool:
        pool.map(_apply_shard, [(lo, min(hi, lo + 4), basis, Mn) for lo, hi in shards[:1]])
        t0 = time.perf_counter()
        chk = pool.map(_apply_shard, [(lo, hi, basis, Mn) for lo, hi in shards])
        dt = time.perf_counter() - t0
    return {"seconds": dt, "nseq": n, "cores": len(shards), "checksum": int(sum(chk))}
