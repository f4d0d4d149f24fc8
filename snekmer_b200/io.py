"""io: file formats of the hot path — drop-in for the parts of ``snekmer.io`` the
vectorize / learn / apply rules use (snekmer/io.py:20-43 ``load_pickle``, :46-96
``load_npz``, :99-122 ``read_kmers``, :167-193 ``define_output_dir``), plus the
FASTA reader that stands in for ``Bio.SeqIO.parse(..., "fasta")`` as the rules use
it (kmerize.smk:90-129: ``f.id``, ``f.seq``, ``len(f.seq)``).
Host-only; the packed (residues, offsets) form is what goes to the device.
"""
from __future__ import annotations

import gzip
import pickle
from ast import literal_eval
from os.path import basename, join, splitext
from typing import Any, Dict, List, Tuple

import numpy as np
import pandas as pd

from .alphabet import ALPHABET_ORDER


def load_pickle(filename: str, mode: str = "rb") -> Any:
    with open(filename, mode) as f:
        return pickle.load(f)


def load_npz(filename: str,
             columns: Dict[str, str] = {"ids": "sequence_id", "seqs": "sequence", "vecs": "sequence_vector"},
             objects: Tuple = ("kmerlist",)) -> Tuple[List, pd.DataFrame]:
    """([kmerlist, ...], DataFrame[filename, sequence_id, sequence, sequence_length, sequence_vector])."""
    data = np.load(filename)
    table: Dict[str, Any] = {"filename": splitext(basename(filename))[0]}
    for key, name in columns.items():
        table[name] = list(data[key])
        if "seq" in key:
            table[f"{name}_length"] = [len(s) for s in data[key]]
    return [data[obj] for obj in objects], pd.DataFrame(table)


def read_kmers(filename: str) -> List[str]:
    """One k-mer per line (basis.txt)."""
    with open(filename) as f:
        return [line.strip() for line in f]


def define_output_dir(alphabet, k: int, nested: bool = False) -> str:
    if isinstance(nested, str):
        nested = literal_eval(nested)
    if not nested:
        return "output"
    name = alphabet if isinstance(alphabet, str) else ALPHABET_ORDER[alphabet]
    return join("output", name, f"k-{k:02}")


# ---------------------------------------------------------------------------
# FASTA -> packed buffer
# ---------------------------------------------------------------------------
def read_fasta(path: str) -> Tuple[List[str], List[str]]:
    """(ids, sequences).  id = header up to the first whitespace, sequence = the record's lines, each
    right-stripped, joined, blanks and CR removed (Bio.SeqIO "fasta" semantics: SimpleFastaParser);
    .gz files are read transparently."""
    opener = gzip.open if str(path).endswith(".gz") else open
    ids: List[str] = []
    seqs: List[str] = []
    parts: List[str] = []
    started = False
    with opener(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if started:
                    seqs.append("".join(parts).replace(" ", "").replace("\r", ""))
                head = line[1:].split(None, 1)
                ids.append(head[0] if head else "")
                parts = []
                started = True
            elif started:
                parts.append(line.rstrip())
    if started:
        seqs.append("".join(parts).replace(" ", "").replace("\r", ""))
    return ids, seqs


def read_fasta_packed(path: str, threads: int = 0, pinned: bool = False):
    """FASTA file -> (ids list[str], residues uint8 [R], offsets int64 [N+1]) with the native multithreaded
    parser (skm_fasta_scan / skm_fasta_pack; same record semantics as read_fasta).  The arrays have the layout
    of engine.SequenceBatch; with pinned=True residues / offsets are views of pinned torch tensors, ready for an
    asynchronous upload."""
    import ctypes as C

    from ._native import check, lib

    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        text = f.read()
    return parse_fasta_bytes(text, threads, pinned)


def parse_fasta_bytes(text: bytes, threads: int = 0, pinned: bool = False):
    import ctypes as C

    from ._native import check, lib

    n = len(text)
    buf = np.frombuffer(text, dtype=np.uint8)
    tp = buf.ctypes.data if n else None
    nseq, nres, idb = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    check(lib().skm_fasta_scan(tp, n, int(threads), C.byref(nseq), C.byref(nres), C.byref(idb)))
    nseq, nres, idb = nseq.value, nres.value, idb.value
    if pinned:
        import torch

        residues = torch.empty(max(nres, 1), dtype=torch.uint8, pin_memory=True).numpy()[:nres]
        offsets = torch.empty(nseq + 1, dtype=torch.int64, pin_memory=True).numpy()
    else:
        residues = np.empty(nres, dtype=np.uint8)
        offsets = np.empty(nseq + 1, dtype=np.int64)
    ids_buf = np.empty(idb, dtype=np.uint8)
    id_off = np.empty(nseq + 1, dtype=np.int64)
    check(lib().skm_fasta_pack(tp, n, int(threads), residues.ctypes.data if nres else None, offsets.ctypes.data,
                               ids_buf.ctypes.data if idb else None, id_off.ctypes.data))
    raw = ids_buf.tobytes()
    bounds = id_off.tolist()
    if raw.isascii():                   # byte offsets are character offsets: one decode, then slices
        txt = raw.decode("ascii")
        ids = [txt[a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    else:
        ids = [raw[a:b].decode("utf-8", "replace") for a, b in zip(bounds[:-1], bounds[1:])]
    return ids, residues, offsets


def pack_sequences(seqs) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(residues uint8[R], starts int64[N+1], raw lengths int64[N]) for a list of strings."""
    lengths = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    residues = np.frombuffer("".join(seqs).encode("latin-1", "replace"), dtype=np.uint8)
    return residues, offsets, lengths
