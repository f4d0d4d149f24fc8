"""vectorize: k-mer vectors — drop-in for ``snekmer.vectorize``.

Same classes, call signatures and results as the reference module
(snekmer/vectorize.py): ``KmerBasis`` (:18-125), ``KmerSet`` (:135-169),
``reduce`` (:173-195), ``make_feature_matrix`` (:201-221), ``KmerVec``
(:224-345).  The per-sequence methods keep their contracts so existing rule
files and pickles keep working; the work itself runs on the GPU through
``snekmer_b200.engine`` (there is no CPU path).  Batch methods (``*_batch``,
``build_basis``) are additions: they are what the rewritten rule bodies in
``snekmer_b200.rules`` call, one launch per FASTA shard instead of one Python
iteration per k-mer.

``KmerVec`` stays picklable (kmerize.smk:141-142 pickles it, search/model/cluster
unpickle it): attributes ``alphabet, k, char_set, vector, basis,
snekmer_version, kmer_set``; no device handle is ever stored on the object.
"""
from __future__ import annotations

import itertools
from typing import Dict, Generator, Iterable, List, Optional, Sequence, Set, Union

import numpy as np

from ._version import __version__
from .alphabet import FULL_ALPHABETS, get_alphabet, get_alphabet_keys
from .utils import check_list


class KmerBasis:
    """An ordered k-mer basis and the change of basis onto it (vectorize.py:18-125)."""

    def __init__(self):
        self.basis = []
        self.basis_order = {}

    def set_basis(self, basis):
        if not check_list(basis):
            raise TypeError("`basis` input must be list or array-like.")
        self.basis = basis
        self.basis_order = dict(enumerate(basis))

    def transform(self, vector, vector_basis):
        """(m, n) vectors over `vector_basis` → (m, p) over the stored basis.

        Output column j is the input column of ``basis[j]`` in `vector_basis`, or a
        zero column when that k-mer is absent (vectorize.py:54-119).  The column
        map is built with one dictionary pass (the reference scans the basis list
        per k-mer); with a CUDA device the gather runs there."""
        if not check_list(vector_basis):
            raise TypeError("`vector_basis` input must be list or array-like.")
        if not isinstance(vector, np.ndarray):
            vector = np.asarray(vector)
        try:
            width = vector.shape[1]
        except IndexError:
            width = len(vector)
        if width != len(vector_basis):
            raise ValueError(
                "Vector and supplied basis shapes must match (vector shape ="
                f" {vector.shape} and len(vector_basis) = {len(vector_basis)}).")
        # last occurrence wins for a k-mer repeated in vector_basis (dict comprehension in the reference)
        where = {kmer: i for i, kmer in enumerate(vector_basis)}
        gather = np.fromiter((where.get(kmer, width) for kmer in self.basis), dtype=np.int64, count=len(self.basis))
        from . import engine

        return engine.gather_columns(vector, gather)


def _generate(alphabet: Set[str], k: int):
    for combo in itertools.product(alphabet, repeat=k):
        yield "".join(combo)


class KmerSet:
    """Holds a k-mer list; enumerates |A|^k k-mers only when none is given (vectorize.py:135-169)."""

    def __init__(self, alphabet: Union[str, int], k: int, kmers: list = None):
        self.alphabet = alphabet
        self.k = k
        self._kmerlist = list(_generate(get_alphabet_keys(alphabet), k)) if kmers is None else kmers

    @property
    def kmers(self):
        return iter(self._kmerlist)


def reduce(sequence: str, alphabet: Union[str, int], mapping: dict = FULL_ALPHABETS) -> str:
    """Trailing '*' removed, then every residue replaced by its reduced symbol
    (unmapped characters stay).  A single string is a host-side ``str.translate``;
    whole shards go through ``engine.reduce_bytes`` on the device."""
    text = str(sequence).rstrip("*")
    table: Dict[str, str] = get_alphabet(alphabet, mapping=mapping)
    return text.translate(str.maketrans(table))


def make_feature_matrix(vecs, min_filter=1, max_filter=1):
    """Ragged per-sequence k-mer lists → (list of 0/1 rows, sorted k-mer list kept when count > min_filter)
    (vectorize.py:201-221).  Small host-side helper kept for API compatibility."""
    flat: List[str] = []
    for v in vecs:
        flat.extend(v)
    kmerlist, counts = np.unique(flat, return_counts=True)
    kmerlist = kmerlist[counts > min_filter]
    rows = []
    for v in vecs:
        row = np.zeros(len(kmerlist))
        row[np.isin(kmerlist, v)] = 1
        rows.append(row)
    return rows, kmerlist


class KmerVec:
    def __init__(self, alphabet: Union[str, int], k: int):
        self.alphabet = alphabet
        self.k = k
        self.char_set = get_alphabet_keys(alphabet)
        self.vector = None
        self.basis = KmerBasis()
        self.snekmer_version = __version__

    def set_kmer_set(self, kmer_set=list()):
        self.kmer_set = KmerSet(self.alphabet, self.k, kmer_set)
        self.basis.set_basis(kmer_set)

    # -- reference-compatible per-sequence API ------------------------------------------
    def _kmer_gen(self, sequence: str) -> Generator[str, None, None]:
        """Valid k-mers of an already reduced string, in order (vectorize.py:239-249)."""
        for kmer in self._valid_kmers([sequence], reduced=True)[0]:
            yield kmer

    @staticmethod
    def _kmer_gen_str(sequence: str, k: int) -> Generator[str, None, None]:
        for n in range(0, len(sequence) - k + 1):
            yield sequence[n:n + k]

    def vectorize(self, sequence: str):
        """The reference method is dead code that raises ``KeyError: 0`` on the first
        k-mer of the set (vectorize.py:282-285); with an empty set it returns {}."""
        vector = {}
        for i, _ in enumerate(self.kmer_set.kmers):
            raise KeyError(i)
        return vector

    def reduce_vectorize(self, sequence: str) -> np.ndarray:
        """Array of the valid reduced k-mers of `sequence`, in order, with repeats
        (vectorize.py:292-328); ``array([], dtype='<U1')`` when there is none."""
        return self._valid_kmers([str(sequence)], reduced=False)[0]

    def harmonize(self, record, kmerlist):
        return self.basis.transform(record, kmerlist)

    # -- batch API (additions) ------------------------------------------------------------
    def _tables(self, reduced: bool):
        from . import engine

        if reduced:
            from . import alphabet as A

            return engine.alphabet_tables_from_symbols(A.symbols(self.alphabet))
        return engine.alphabet_tables(self.alphabet)

    def _valid_kmers(self, sequences: Sequence[str], reduced: bool) -> List[np.ndarray]:
        from . import engine

        batch = engine.SequenceBatch.from_strings(sequences)
        tab = self._tables(reduced)
        codes = engine.encode_windows(batch, tab, self.k).cpu().numpy()
        if codes.dtype == np.int32:
            codes, none = codes.view(np.uint32), np.uint32(0xFFFFFFFF)
        else:
            codes, none = codes.view(np.uint64), np.uint64(0xFFFFFFFFFFFFFFFF)
        out = []
        off = batch.offsets_host
        for i in range(batch.n):
            c = codes[off[i]:off[i + 1]]
            c = c[c != none]
            out.append(engine.decode_kmers(c.astype(np.uint64), tab.symbols, self.k) if c.size
                       else np.array([], dtype="<U1"))
        return out

    def reduce_vectorize_batch(self, sequences: Iterable[str]) -> List[np.ndarray]:
        """``[self.reduce_vectorize(s) for s in sequences]`` in one launch."""
        return self._valid_kmers([str(s) for s in sequences], reduced=False)

    def build_basis(self, sequences: Iterable[str], min_filter: int = 0) -> np.ndarray:
        """The k-mer basis of a set of sequences as the vectorize rule builds it
        (kmerize.smk:89-104): first-occurrence order, total count > min_filter.
        Also installs it with ``set_kmer_set``."""
        from . import engine

        batch = engine.SequenceBatch.from_strings([str(s) for s in sequences])
        basis = engine.build_basis(batch, self.alphabet, self.k, min_filter)
        kmers = basis.kmers()
        self.set_kmer_set(kmers)
        return kmers

    def count_batch(self, sequences: Iterable[str], kmerlist: Optional[Sequence[str]] = None) -> np.ndarray:
        """int32 [N, K] counts of every sequence over `kmerlist` (default: the installed
        k-mer set) — what learn.smk:359-383 / apply.smk:195-206 compute per sequence."""
        from . import engine

        kmerlist = list(self.kmer_set.kmers) if kmerlist is None else list(kmerlist)
        batch = engine.SequenceBatch.from_strings([str(s) for s in sequences])
        return engine.count_over_kmers(batch, self.alphabet, self.k, kmerlist).cpu().numpy()
