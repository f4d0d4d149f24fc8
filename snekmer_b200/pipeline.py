"""pipeline: host-buffer entry points (what a caller with data in host RAM uses).

``vectorize_host`` is the whole vectorize rule body (kmerize.smk:67-129) for a packed FASTA shard held in (pinned)
host memory: residues go to HBM in chunks on a copy stream while pass 1 (basis) already runs on the chunks that
have landed; after the basis is finalised, pass 2 (dense counts) runs chunk by chunk and each finished block of rows
is copied back to the caller's host buffer on a second copy stream, double-buffered.  PCIe is full duplex, so the
input upload, both kernels and the result download overlap.  The download is what bounds it (a 1 M x 1000 int32
matrix is 4 GB), so the rows can travel in a compact form:

  transport   bytes per count   content
  "int32"     4                 the count matrix as the kernels produce it
  "uint16"    2                 lossless while no sequence has more than 65,535 residues (checked by the library)
  "uint8"     1 (+ escapes)     counts < 255 as bytes; the few larger ones as an escape list (row, col, count): lossless
  "bits"      1/8               the presence matrix, bit-packed — what the reference's vectorize rule itself hands on
                                (`vecs`, kmerize.smk:112-120; its learn / apply rules re-count from the sequences)

``apply_host`` is the apply rule's device part for host residues: counts never leave HBM, they feed the tcgen05
scoring directly, and 20 bytes per query (top-1, top-2 ids and their cosine scores) come back.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import engine as E
from ._native import check, lib

TRANSPORTS = ("int32", "uint16", "uint8", "bits")


def _chunks_by_residues(offsets: np.ndarray, n_chunks: int):
    n = len(offsets) - 1
    if n <= 0:
        return []
    n_chunks = max(1, min(n_chunks, n))
    targets = offsets[0] + (offsets[-1] - offsets[0]) * np.arange(1, n_chunks) / n_chunks
    cuts = np.unique(np.concatenate([[0], np.searchsorted(offsets[:-1], targets, side="left"), [n]])).astype(np.int64)
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]


# ---------------------------------------------------------------------------
# compact transports
# ---------------------------------------------------------------------------
def pack_counts_u8(counts: torch.Tensor, out: Optional[torch.Tensor] = None, esc_capacity: Optional[int] = None):
    """Device count matrix (int32 / uint16 [N, K]) -> (uint8 [N, K], esc_row, esc_col, esc_val int32 [n_esc]).
    Entries >= 255 read 255 in the byte matrix and are listed in the escape arrays (unordered)."""
    dev = E._require_cuda(counts.device)
    counts = counts.contiguous()
    rows, cols = counts.shape
    bits = {torch.int32: 32, torch.uint16: 16, torch.int16: 16}[counts.dtype]
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.uint8, device=dev)
    cap = max(1024, rows // 4) if esc_capacity is None else int(esc_capacity)
    while True:
        esc = torch.empty((3, cap), dtype=torch.int32, device=dev)
        dn = torch.zeros(1, dtype=torch.int64, device=dev)
        check(lib().skm_pack_counts_u8(E._ptr(counts), rows, cols, bits, E._ptr(out), E._ptr(esc[0]), E._ptr(esc[1]), E._ptr(esc[2]),
                                       cap, E._ptr(dn), E._stream()))
        n = int(dn.item())
        if n <= cap:
            return out, esc[0, :n], esc[1, :n], esc[2, :n]
        cap = n                                     # rare: a shard of low-complexity sequences; once more with room


def unpack_counts_u8(u8: np.ndarray, esc_row: np.ndarray, esc_col: np.ndarray, esc_val: np.ndarray) -> np.ndarray:
    """Host inverse of pack_counts_u8: int32 [N, K]."""
    out = u8.astype(np.int32)
    out[esc_row, esc_col] = esc_val
    return out


def pack_presence_bits(counts: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device count matrix -> uint8 [N, ceil(K / 8)]: bit (c & 7) of byte c >> 3 is counts[r, c] > 0
    (numpy.unpackbits(x, axis=1, bitorder="little")[:, :K])."""
    dev = E._require_cuda(counts.device)
    counts = counts.contiguous()
    rows, cols = counts.shape
    bits = {torch.int32: 32, torch.uint16: 16, torch.int16: 16}[counts.dtype]
    if out is None:
        out = torch.empty((rows, (cols + 7) // 8), dtype=torch.uint8, device=dev)
    check(lib().skm_pack_presence_bits(E._ptr(counts), rows, cols, bits, E._ptr(out), E._stream()))
    return out


@dataclass
class HostVectors:
    """Result of vectorize_host: the basis (device tables + host k-mers on demand) and the rows in host memory."""
    basis: E.Basis
    transport: str
    n: int
    K: int
    data: torch.Tensor                      # host tensor in the transport's layout
    escapes: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]] = None     # uint8 transport: (row, col, count)

    def counts(self) -> np.ndarray:
        """int32 [N, K] (not available for the presence-bit transport)."""
        if self.transport == "bits":
            raise ValueError("the presence-bit transport carries no counts")
        a = self.data.numpy()
        if self.transport == "uint8":
            return unpack_counts_u8(a, *self.escapes)
        return a.astype(np.int32, copy=False) if self.transport == "int32" else a.astype(np.int32)

    def presence(self) -> np.ndarray:
        """bool [N, K] — `vecs` of the reference's .npz is this as float64."""
        if self.transport == "bits":
            return np.unpackbits(self.data.numpy(), axis=1, bitorder="little")[:, :self.K].astype(bool)
        return self.counts() > 0


def _enqueue_uploads(residues: torch.Tensor, offsets: np.ndarray, chunks, dev, s_in, main):
    """Residues -> HBM chunk by chunk on the copy stream, ALL enqueued up front.  Returns (batches, events): batch i may
    be used on a stream that has waited for event i."""
    nres = int(offsets[-1])
    d_res = torch.empty(max(nres, 1) + 16, dtype=torch.uint8, device=dev)
    d_off = torch.from_numpy(offsets).to(dev, non_blocking=True)
    s_in.wait_stream(main)
    batches, events = [], []
    for lo, hi in chunks:
        r0, r1 = int(offsets[lo]), int(offsets[hi])
        with torch.cuda.stream(s_in):
            if r1 > r0:
                d_res[r0:r1].copy_(residues[r0:r1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_in)
        b = E.SequenceBatch(d_res, d_off[lo:hi + 1], offsets[lo:hi + 1])
        b.nres = r1                      # kernels may read the buffer up to the end of this chunk
        batches.append(b)
        events.append(ev)
    return batches, events


def vectorize_host(residues: torch.Tensor, offsets: np.ndarray, alphabet, k: int, min_filter: int = 0,
                   out: Optional[torch.Tensor] = None, transport: str = "int32", n_chunks: int = 16,
                   device=None, dtype: Optional[torch.dtype] = None) -> HostVectors:
    """Host residues (uint8 tensor, ideally pinned) + host offsets -> HostVectors (basis + rows in host memory).

    `out` may be a pre-allocated pinned host tensor in the transport's layout ([N, K] int32 / uint16 / uint8, or
    [N, ceil(K/8)] uint8 for "bits") to receive the rows.  `dtype` (torch.int32 / torch.uint16) is the round-1
    spelling of transport."""
    if dtype is not None:
        transport = {torch.int32: "int32", torch.uint16: "uint16"}[dtype]
    if transport not in TRANSPORTS:
        raise ValueError(f"transport must be one of {TRANSPORTS}")
    dev = E._require_cuda(device)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    tab = E.alphabet_tables(alphabet, dev)
    S = E.code_space(tab.nsym, k)
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    chunks = _chunks_by_residues(offsets, n_chunks)
    order_only = min_filter <= 0 and E.order_only_supported(S)
    batches, up_events = _enqueue_uploads(residues, offsets, chunks, dev, s_in, main)
    if order_only:
        # The basis order needs first positions only, and they are final once every code of the space has one.  The walk
        # follows the upload chunk by chunk and the 4-byte "saturated" flag is read back after each chunk: as soon as
        # it is set the basis is finalised and the count pass starts on the chunks already in HBM while the rest of the
        # upload is still in flight (PCIe is full duplex: the upload then hides behind the download).
        first = torch.full((S,), -1, dtype=torch.int64, device=dev)
        state = torch.zeros(4, dtype=torch.int32, device=dev)
        for b, ev in zip(batches, up_events):
            main.wait_event(ev)
            E.basis_first_progressive(b, alphabet, k, first, state, 0)
            if int(state[0].item()):
                break
        basis = E.basis_finalize(alphabet, k, None, first, 0)           # reads K back (one 8-byte sync)
    else:
        count, first = E.basis_tables(S, dev)
        for b, ev in zip(batches, up_events):
            main.wait_event(ev)
            E.basis_accumulate(b, alphabet, k, count, first, 0)
        basis = E.basis_finalize(alphabet, k, count, first, min_filter)
    K = basis.K
    host_dtype = {"int32": torch.int32, "uint16": torch.uint16, "uint8": torch.uint8, "bits": torch.uint8}[transport]
    width = (K + 7) // 8 if transport == "bits" else K
    if out is None:
        out = torch.empty((n, width), dtype=host_dtype, pin_memory=True)
    assert tuple(out.shape) == (n, width) and out.dtype == host_dtype and out.is_contiguous()
    rows_max = max(hi - lo for lo, hi in chunks) if chunks else 0
    cnt_dtype = torch.uint16 if transport == "uint16" else torch.int32
    packed = transport in ("uint8", "bits")
    bufs = [torch.empty((rows_max, K), dtype=cnt_dtype, device=dev) for _ in range(1 if packed else 2)]
    pbufs = [torch.empty((rows_max, width), dtype=torch.uint8, device=dev) for _ in range(2)] if packed else None
    esc_cap = max(4096, rows_max // 2)
    esc_bufs = [torch.empty((3, esc_cap), dtype=torch.int32, device=dev) for _ in range(2)] if transport == "uint8" else None
    esc_n = torch.zeros(len(chunks) + 1, dtype=torch.int64, device=dev) if transport == "uint8" else None
    esc_host = [torch.empty((3, esc_cap), dtype=torch.int32, pin_memory=True) for _ in chunks] if transport == "uint8" else None
    free_ev = [None, None]
    for i, ((lo, hi), b) in enumerate(zip(chunks, batches)):
        slot = i & 1
        main.wait_event(up_events[i])
        if free_ev[slot] is not None:
            main.wait_event(free_ev[slot])
        cbuf = bufs[0 if packed else slot][: hi - lo]
        E.count_dense(b, alphabet, k, basis, dtype=cnt_dtype, out=cbuf)
        if transport == "uint8":
            src = pbufs[slot][: hi - lo]
            e = esc_bufs[slot]
            check(lib().skm_pack_counts_u8(E._ptr(cbuf), hi - lo, K, 32, E._ptr(src), E._ptr(e[0]), E._ptr(e[1]), E._ptr(e[2]),
                                           esc_cap, E._ptr(esc_n[i:]), E._stream()))
        elif transport == "bits":
            src = pack_presence_bits(cbuf, out=pbufs[slot][: hi - lo])
        else:
            src = cbuf
        done = torch.cuda.Event()
        done.record(main)
        s_out.wait_event(done)
        with torch.cuda.stream(s_out):
            out[lo:hi].copy_(src, non_blocking=True)
            if transport == "uint8":
                esc_host[i].copy_(esc_bufs[slot], non_blocking=True)
            fe = torch.cuda.Event()
            fe.record(s_out)
        free_ev[slot] = fe
    main.wait_stream(s_out)
    escapes = None
    if transport == "uint8":
        counts_n = esc_n.cpu().numpy()                  # synchronises: everything above has landed
        rows_l: List[np.ndarray] = []
        cols_l: List[np.ndarray] = []
        vals_l: List[np.ndarray] = []
        for i, (lo, hi) in enumerate(chunks):
            m = int(counts_n[i])
            if m > esc_cap:                             # rare: redo this chunk's escape list with room for all of it
                cbuf = E.count_dense(batches[i], alphabet, k, basis, dtype=torch.int32)
                _, er, ec, evv = pack_counts_u8(cbuf, esc_capacity=m)
                e = torch.stack([er, ec, evv]).cpu().numpy()
            else:
                e = esc_host[i].numpy()[:, :m]
            rows_l.append(e[0].astype(np.int64) + lo)
            cols_l.append(e[1].astype(np.int64))
            vals_l.append(e[2].astype(np.int32))
        cat = (lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt))
        escapes = (cat(rows_l, np.int64), cat(cols_l, np.int64), cat(vals_l, np.int32))
    return HostVectors(basis, transport, n, K, out, escapes)


@dataclass
class HostScores:
    """apply_host result in host memory (20 bytes per query)."""
    top1: np.ndarray        # int32 [Q]
    top2: np.ndarray        # int32 [Q]
    score1: np.ndarray      # float64 [Q]
    score2: np.ndarray      # float64 [Q]


def apply_host(residues: torch.Tensor, offsets: np.ndarray, alphabet, k: int, prepared: E.PreparedAnnotations,
               basis: Optional[E.Basis] = None, n_chunks: int = 4, device=None,
               out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]] = None) -> HostScores:
    """Host query residues -> host top-2 annotations and cosine scores against a prepared annotation matrix
    (apply.smk:188-206 counts + :278-335 scoring).  Chunks of queries are uploaded on a copy stream while the previous
    chunk is counted and scored; the count rows stay in HBM.  `basis`: columns of the prepared matrix (None = the
    identity basis: column = code).  `out`: pinned host tensors (top1, top2 int32 [Q]; score1, score2 float64 [Q])."""
    dev = E._require_cuda(device)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    chunks = _chunks_by_residues(offsets, n_chunks)
    if out is None:
        out = (torch.empty(n, dtype=torch.int32, pin_memory=True), torch.empty(n, dtype=torch.int32, pin_memory=True),
               torch.empty(n, dtype=torch.float64, pin_memory=True), torch.empty(n, dtype=torch.float64, pin_memory=True))
    batches, up_events = _enqueue_uploads(residues, offsets, chunks, dev, s_in, main)
    rows_max = max(hi - lo for lo, hi in chunks) if chunks else 0
    K = prepared.K
    Qbuf = torch.empty((rows_max, K), dtype=torch.int32, device=dev)
    for (lo, hi), b, ev in zip(chunks, batches, up_events):
        main.wait_event(ev)
        Q = E.count_dense(b, alphabet, k, basis, out=Qbuf[: hi - lo])
        r = E.apply_tc(Q, prepared, E.row_norm2(Q))
        if r is None:
            raise E.SkmError(-3, "apply_host: a query count exceeds 255 and the prepared matrix does not carry M")
        done = torch.cuda.Event()
        done.record(main)
        s_out.wait_event(done)
        with torch.cuda.stream(s_out):
            for dst, src in zip(out, (r.top1, r.top2, r.score1, r.score2)):
                dst[lo:hi].copy_(src, non_blocking=True)
                src.record_stream(s_out)
    main.wait_stream(s_out)
    torch.cuda.current_stream(dev).synchronize()
    return HostScores(out[0].numpy(), out[1].numpy(), out[2].numpy(), out[3].numpy())


@dataclass
class HostCOO:
    """learn_host result in host memory: the annotation x k-mer count matrix as a COO list sorted by key = ann * S + code.
    `packed` (uint64 [nnz], key << count_bits | count) is what crossed PCIe — 8 bytes per entry; keys() / vals() unpack."""
    S: int
    n_ann: int
    count_bits: int
    packed: Optional[np.ndarray] = None     # uint64 [nnz]
    raw: Optional[Tuple[np.ndarray, np.ndarray]] = None     # (keys int64, vals int64) when a count did not fit count_bits

    @property
    def nnz(self) -> int:
        return int(self.packed.size if self.packed is not None else self.raw[0].size)

    def keys(self) -> np.ndarray:
        return (self.packed >> np.uint64(self.count_bits)).astype(np.int64) if self.packed is not None else self.raw[0]

    def vals(self) -> np.ndarray:
        if self.packed is None:
            return self.raw[1]
        return (self.packed & np.uint64((1 << self.count_bits) - 1)).astype(np.int64)


def learn_host(residues: torch.Tensor, offsets: np.ndarray, ann_id, alphabet, k: int, n_ann: int,
               out: Optional[torch.Tensor] = None, device=None) -> HostCOO:
    """Host residues + annotation ids -> the learned count matrix in host memory (learn.smk:306-326, 385-408 for one
    shard).  The matrix is built on the device (engine.learn_sparse_hybrid) and comes back in the packed exchange format
    of the multi-GPU fan-in (skm_coo_pack: 8 bytes per entry instead of 16).  `out`: optional pinned int64 tensor with
    room for the entries (at most one per residue)."""
    dev = E._require_cuda(device)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    batch = E.SequenceBatch.from_packed(residues.numpy() if isinstance(residues, torch.Tensor) else residues, offsets, dev,
                                        pinned=isinstance(residues, torch.Tensor) and residues.is_pinned())
    ann = ann_id if isinstance(ann_id, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(ann_id, dtype=np.int32))
    keys, vals = E.learn_sparse_hybrid(batch, alphabet, k, ann.to(dev, non_blocking=True), n_ann)
    S = E.code_space(E.alphabet_tables(alphabet, dev).nsym, k)
    count_bits = 64 - max(int(n_ann) * int(S) - 1, 1).bit_length()
    m = int(keys.numel())
    if count_bits >= 16:
        packed, bad = E.coo_pack(keys, vals, count_bits)
        if not bad:
            if out is None or out.numel() < m:
                out = torch.empty(max(m, 1), dtype=torch.int64, pin_memory=True)
            out[:m].copy_(packed, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return HostCOO(S, int(n_ann), count_bits, packed=out[:m].numpy().view(np.uint64))
    hk, hv = keys.cpu().numpy(), vals.cpu().numpy()
    return HostCOO(S, int(n_ann), count_bits, raw=(hk, hv))
