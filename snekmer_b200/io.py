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
    """(ids, sequences).  id = header up to the first whitespace, sequence = the
    record's lines joined (Bio.SeqIO "fasta" semantics); .gz files are read transparently."""
    opener = gzip.open if str(path).endswith(".gz") else open
    ids: List[str] = []
    seqs: List[str] = []
    parts: List[str] = []
    started = False
    with opener(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if started:
                    seqs.append("".join(parts))
                head = line[1:].split(None, 1)
                ids.append(head[0] if head else "")
                parts = []
                started = True
            elif started:
                parts.append(line.strip())
    if started:
        seqs.append("".join(parts))
    return ids, seqs


def pack_sequences(seqs) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(residues uint8[R], starts int64[N+1], raw lengths int64[N]) for a list of strings."""
    lengths = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    residues = np.frombuffer("".join(seqs).encode("latin-1", "replace"), dtype=np.uint8)
    return residues, offsets, lengths
