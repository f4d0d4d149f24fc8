"""rules_sparse: the learn / merge / apply rule bodies for bases that do not fit dense matrices.

``rules.py`` mirrors the reference one to one: a dense ``[annotations, K]`` matrix in memory and wide CSV files
(learn.smk:306-357, apply.smk:147-353).  At the benchmark shapes (C3 / C4: 20 k - 50 k annotations x 1.68 M k-mers) that
matrix is 100+ GB dense and its CSV far larger, so the same three steps are offered on the sparse device paths with the
binary ``.skmc`` side-car (sidecar.py) as the file between them:

  learn_counts_sparse   Library (learn.smk:306-408)      -> COO matrix + Totals on the device
  merge_counts_sparse   Merge (learn.smk:467-494)        -> outer join on annotations, sum (skm_coo_merge)
  apply_counts_sparse   KmerCompare (apply.smk:224-342)  -> SpMM cosine, top-2, delta, Confidence

Row / column conventions are the reference's: annotation rows in order of first appearance, ``Totals`` over ALL
sequences, a repeated sequence id keeps its first position and last value, the query norm runs over all valid windows of
the query (its own k-mer list), ties go to the lowest annotation index.  ``sidecar.export_counts_csv`` /
``export_totals_csv`` turn a side-car back into the reference's CSV when the matrix is small enough to be one.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import engine as E
from . import sidecar
from .rules import _ACCESSION, ScoreResult, _unique_last, _write_arrow_csv, delta_and_confidence, read_confidence_csv


@dataclass
class SparseCounts:
    """A learned count matrix on the device: sorted COO keys ``ann * S + code`` over the reduced alphabet ``symbols``."""
    annotations: List[str]
    seq_count: np.ndarray           # int64 [A]
    symbols: str
    k: int
    keys: torch.Tensor              # int64 [nnz], sorted
    vals: torch.Tensor              # int64 [nnz]
    totals: torch.Tensor            # int64 [S]: occurrences over ALL sequences
    total_seqs: int

    @property
    def S(self) -> int:
        return len(self.symbols) ** self.k


def annotation_ids(ids: Sequence[str], seq_annot: Dict[str, str]):
    """(ann_id int32 [N], annotations in first-appearance order, sequences per annotation, distinct ids): the dict
    semantics of learn.smk:316-326 — a repeated id keeps its first position but only its LAST record counts."""
    last = {}
    for i, sid in enumerate(ids):
        last[sid] = i
    index: Dict[str, int] = {}
    ann_id = np.full(len(ids), -1, dtype=np.int32)
    seq_count: List[int] = []
    for sid, i in last.items():
        acc = _ACCESSION.findall(sid)[0]              # IndexError without two pipes, like learn.smk:322
        if acc not in seq_annot:
            continue
        a = seq_annot[acc]
        if a not in index:
            index[a] = len(index)
            seq_count.append(0)
        ann_id[i] = index[a]
        seq_count[index[a]] += 1
    return ann_id, list(index), np.array(seq_count, dtype=np.int64), len(last)


def learn_counts_sparse(ids: Sequence[str], reduced_seqs: Sequence[str], symbols: str, k: int,
                        seq_annot: Dict[str, str]) -> SparseCounts:
    """Library.generate_kmer_counts + filter_and_construct for one file, sparse.  NOTE the reference adds EVERY record
    to the totals but only the last record of a repeated id to its annotation row; so does this (records that lost to a
    later duplicate are counted as unannotated)."""
    tab = E.alphabet_tables_from_symbols(symbols)
    batch = E.SequenceBatch.from_strings([str(s) for s in reduced_seqs])
    ann_id, annotations, seq_count, total = annotation_ids(ids, seq_annot)
    keys, vals, totals = E.learn_sparse_with_totals(batch, tab, int(k), torch.from_numpy(ann_id), len(annotations))
    return SparseCounts(annotations, seq_count, tab.symbols, int(k), keys, vals, totals, total)


def write_counts_sidecar(path: str, sc: SparseCounts, kmers: Optional[Sequence[str]] = None) -> None:
    """``.skmc`` of a learned matrix: rows ``Totals`` + annotations, columns = `kmers` (default: the k-mers that occur, in
    code order; pass the vectorize rule's kmerlist to get the reference's column order), CSR values."""
    S = sc.S
    totals = sc.totals.cpu().numpy()
    if kmers is None:
        codes = np.flatnonzero(totals).astype(np.uint64)
        kmers = list(E.decode_kmers(codes, sc.symbols, sc.k))
        col_of_code = np.full(S, -1, dtype=np.int64)
        col_of_code[codes.astype(np.int64)] = np.arange(len(codes))
    else:
        kmers = [str(x) for x in kmers]
        codes, ok = E.encode_kmers(kmers, sc.symbols, sc.k)
        col_of_code = np.full(S, -1, dtype=np.int64)
        col_of_code[codes[ok].astype(np.int64)] = np.flatnonzero(ok)
    keys = sc.keys.cpu().numpy()
    vals = sc.vals.cpu().numpy()
    ann, code = keys // S, keys % S
    col = col_of_code[code]
    keep = col >= 0
    ann, col, vals = ann[keep], col[keep], vals[keep]
    order = np.lexsort((col, ann))                                   # rows by annotation, columns ascending inside a row
    ann, col, vals = ann[order], col[order], vals[order]
    t_cols = np.flatnonzero(col_of_code >= 0)
    t_col = col_of_code[t_cols]
    t_order = np.argsort(t_col, kind="stable")
    t_vals = totals[t_cols][t_order]
    nzt = t_vals != 0
    A = len(sc.annotations)
    rowptr = np.zeros(A + 2, dtype=np.int64)
    rowptr[1] = int(nzt.sum())
    np.cumsum(np.bincount(ann, minlength=A), out=rowptr[2:])
    rowptr[2:] += rowptr[1]
    row_sum = np.zeros(A, dtype=np.int64)
    np.add.at(row_sum, ann, vals)
    sidecar.write_counts_csr(path, ["Totals"] + list(sc.annotations), kmers,
                             np.concatenate([[sc.total_seqs], sc.seq_count]), np.concatenate([[int(t_vals.sum())], row_sum]),
                             rowptr, np.concatenate([t_col[t_order][nzt], col]).astype(np.int32), np.concatenate([t_vals[nzt], vals]))


def read_counts_sidecar(path: str, symbols: Optional[str] = None) -> SparseCounts:
    """``.skmc`` -> device COO.  The reduced alphabet is the set of characters of the k-mer columns unless given."""
    d = sidecar.read_counts_csr(path)
    kmers = d["kmers"]
    if not d["rows"] or d["rows"][0] != "Totals":
        raise ValueError(f"{path}: a counts table starts with its Totals row")
    k = len(kmers[0]) if kmers else 1
    symbols = "".join(sorted(set("".join(kmers)))) if symbols is None else "".join(sorted(set(symbols)))
    S = len(symbols) ** k
    codes, ok = E.encode_kmers(kmers, symbols, k)
    dev = E._require_cuda()
    rowptr, cols, vals = d["rowptr"], d["cols"].astype(np.int64), d["vals"]
    rows = np.repeat(np.arange(len(d["rows"])), np.diff(rowptr))
    good = ok[cols]
    totals = np.zeros(S, dtype=np.int64)
    t = (rows == 0) & good
    totals[codes[cols[t]].astype(np.int64)] = vals[t]
    m = (rows > 0) & good
    keys = (rows[m] - 1).astype(np.int64) * S + codes[cols[m]].astype(np.int64)
    order = np.argsort(keys, kind="stable")
    return SparseCounts(list(d["rows"][1:]), d["seq_count"][1:].astype(np.int64), symbols, k,
                        torch.from_numpy(keys[order]).to(dev), torch.from_numpy(vals[m][order].astype(np.int64)).to(dev),
                        torch.from_numpy(totals).to(dev), int(d["seq_count"][0]))


def merge_counts_sparse(parts: Sequence[SparseCounts]) -> SparseCounts:
    """Merge.merge_dataframes (learn.smk:467-494): outer join on the annotation rows (first appearance), element-wise sum."""
    assert parts and all(p.symbols == parts[0].symbols and p.k == parts[0].k for p in parts), "same reduced alphabet and k"
    S = parts[0].S
    index: Dict[str, int] = {}
    for p in parts:
        for a in p.annotations:
            index.setdefault(a, len(index))
    seq_count = np.zeros(len(index), dtype=np.int64)
    ks, vs = [], []
    totals = torch.zeros_like(parts[0].totals)
    total_seqs = 0
    for p in parts:
        remap = torch.tensor([index[a] for a in p.annotations] + [0], dtype=torch.int64, device=p.keys.device)
        ks.append(remap[p.keys // S] * S + p.keys % S)
        vs.append(p.vals)
        np.add.at(seq_count, [index[a] for a in p.annotations], p.seq_count)
        totals += p.totals
        total_seqs += p.total_seqs
    keys, vals = E.coo_merge(torch.cat(ks), torch.cat(vs), key_bound=len(index) * S)
    return SparseCounts(list(index), seq_count, parts[0].symbols, parts[0].k, keys, vals, totals, total_seqs)


def restrict_csr_to_kmers(rowptr: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, symbols: str, k: int,
                          kmers: Sequence[str]):
    """Drop the CSR entries (codes) that are not in `kmers` — the query file's own kmerlist: the reference takes the
    query norm over that list only (apply.smk:262-289), which matters when the vectorize step ran with min_filter > 0 or
    a basis.txt."""
    S = len(symbols) ** int(k)
    codes, ok = E.encode_kmers([str(x) for x in kmers], symbols, int(k))
    member = torch.zeros(S, dtype=torch.bool, device=cols.device)
    member[torch.from_numpy(codes[ok].astype(np.int64)).to(cols.device)] = True
    keep = member[cols.to(torch.int64)]
    csum = torch.zeros(keep.numel() + 1, dtype=torch.int64, device=cols.device)
    torch.cumsum(keep, 0, out=csum[1:])
    return csum[rowptr], cols[keep].contiguous(), vals[keep].contiguous()


def apply_counts_sparse(ids: Sequence[str], reduced_seqs: Sequence[str], sc: SparseCounts, confidence_csv: Optional[str] = None,
                        out_summary: Optional[str] = None, tile: Optional[int] = None,
                        query_kmers: Optional[Sequence[str]] = None) -> ScoreResult:
    """KmerCompare of the apply workflow (apply.smk:224-342) as SpMM: cosine of every query against every annotation
    row, top-2, delta, Confidence; writes kmer-summary CSV when asked.  A repeated id keeps its last record.
    query_kmers: the query file's kmerlist (from its .npz / side-car).  The reference's query norm and dot run over that
    list; without it every valid window of the query counts — the same thing when the vectorize step ran with
    min_filter = 0 and no basis file, which is the only case the two differ in.  The annotation tile follows the
    accumulator width the data needs (engine.apply_sparse_tiled)."""
    uniq, pick = _unique_last(list(ids))
    seqs = [str(reduced_seqs[int(i)]) for i in pick]
    tab = E.alphabet_tables_from_symbols(sc.symbols)
    batch = E.SequenceBatch.from_strings(seqs)
    rowptr, cols, cvals = E.count_csr(batch, tab, sc.k, None)
    if query_kmers is not None:
        rowptr, cols, cvals = restrict_csr_to_kmers(rowptr, cols, cvals, tab.symbols, sc.k, query_kmers)
    A = len(sc.annotations)
    r = E.apply_sparse_tiled(rowptr, cols, cvals, sc.keys, sc.vals, sc.S, A, tile)
    top1, top2 = r.top1.cpu().numpy(), r.top2.cpu().numpy()
    s1, s2 = r.score1.cpu().numpy(), r.score2.cpu().numpy()
    if out_summary:
        conf = read_confidence_csv(confidence_csv) if confidence_csv else {}
        delta, confidence = delta_and_confidence(r, conf, A)
        import pyarrow as pa

        _write_arrow_csv(out_summary, {"index": uniq, "Prediction": [str(sc.annotations[i]) for i in top1], "Score": s1,
                                       "delta": delta, "Confidence": pa.array(confidence, from_pandas=True)})
    return ScoreResult(uniq, list(sc.annotations), top1, top2, s1, s2, None)
