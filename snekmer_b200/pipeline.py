"""pipeline: host-buffer entry points (what a caller with data in host RAM uses).

``vectorize_host`` is the whole vectorize rule body (kmerize.smk:67-129) for a
packed FASTA shard held in (pinned) host memory: residues go to HBM in chunks on
a copy stream while pass 1 (basis accumulation) already runs on the chunks that
have landed; after the basis is finalised, pass 2 (dense counts) runs chunk by
chunk and each finished block of rows is copied back to the caller's host
buffer on a second copy stream, double-buffered.  PCIe is full duplex, so the
input upload, both kernels and the result download overlap.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import engine as E


def _chunks_by_residues(offsets: np.ndarray, n_chunks: int):
    n = len(offsets) - 1
    if n <= 0:
        return []
    n_chunks = max(1, min(n_chunks, n))
    targets = offsets[0] + (offsets[-1] - offsets[0]) * np.arange(1, n_chunks) / n_chunks
    cuts = np.unique(np.concatenate([[0], np.searchsorted(offsets[:-1], targets, side="left"), [n]])).astype(np.int64)
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]


def vectorize_host(residues: torch.Tensor, offsets: np.ndarray, alphabet, k: int, min_filter: int = 0,
                   out: Optional[torch.Tensor] = None, dtype: torch.dtype = torch.int32, n_chunks: int = 16,
                   device=None) -> Tuple[E.Basis, torch.Tensor]:
    """Host residues (uint8 tensor, ideally pinned) + host offsets → (basis, host count matrix [N, K]).

    `out` may be a pre-allocated pinned host tensor [N, >=K... exactly K] to receive the counts."""
    dev = E._require_cuda(device)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    nres = int(offsets[-1])
    tab = E.alphabet_tables(alphabet, dev)
    S = E.code_space(tab.nsym, k)
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_res = torch.empty(max(nres, 1) + 16, dtype=torch.uint8, device=dev)
    d_off = torch.from_numpy(offsets).to(dev, non_blocking=True)
    count, first = E.basis_tables(S, dev)
    chunks = _chunks_by_residues(offsets, n_chunks)
    s_in.wait_stream(main)
    batches = []
    for lo, hi in chunks:
        r0, r1 = int(offsets[lo]), int(offsets[hi])
        with torch.cuda.stream(s_in):
            if r1 > r0:
                d_res[r0:r1].copy_(residues[r0:r1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_in)
        main.wait_event(ev)
        b = E.SequenceBatch(d_res, d_off[lo:hi + 1], offsets[lo:hi + 1])
        b.nres = r1                      # kernels may read the buffer up to the end of this chunk
        batches.append(b)
        E.basis_accumulate(b, alphabet, k, count, first, 0)
    basis = E.basis_finalize(alphabet, k, count, first, min_filter)      # reads K back (one 8-byte sync)
    K = basis.K
    if out is None:
        out = torch.empty((n, K), dtype=dtype, pin_memory=True)
    assert tuple(out.shape) == (n, K) and out.dtype == dtype and out.is_contiguous()
    rows_max = max(hi - lo for lo, hi in chunks) if chunks else 0
    bufs = [torch.empty((rows_max, K), dtype=dtype, device=dev) for _ in range(2)]
    free_ev = [None, None]
    for i, ((lo, hi), b) in enumerate(zip(chunks, batches)):
        buf = bufs[i & 1][: hi - lo]
        if free_ev[i & 1] is not None:
            main.wait_event(free_ev[i & 1])
        E.count_dense(b, alphabet, k, basis, dtype=dtype, out=buf)
        done = torch.cuda.Event()
        done.record(main)
        s_out.wait_event(done)
        with torch.cuda.stream(s_out):
            out[lo:hi].copy_(buf, non_blocking=True)
            fe = torch.cuda.Event()
            fe.record(s_out)
        free_ev[i & 1] = fe
    main.wait_stream(s_out)
    return basis, out
