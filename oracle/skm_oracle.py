"""CPU oracle for the Snekmer vectorize + learn/apply hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``snekmer_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or the
reported CPU baseline.

It is a numpy/pure-Python *restatement* of the reference's algorithm
(PNNL-CompBio/Snekmer 1.3.0, pure Python), written from its behaviour; every
function cites the reference file:line it follows.  Parity pinning: the
reference ships no golden vectors (its CI only checks exit codes,
``.github/workflows/action.yml:68-124``), so this oracle is pinned against
outputs of the *reference itself run in the build container*
(``tests/golden/make_golden.py`` → ``tests/golden/*``; checked by
``tests/test_oracle_golden.py`` and, where ``/root/reference`` exists, live by
``tests/test_reference_live.py``).  The cosine arithmetic lives in scikit-learn
(unpinned in the reference's requirements.txt:8); it is restated here as
``x/||x|| . y/||y||`` in float64 with zero-norm rows left at zero and pinned
against scikit-learn 1.9.0.

Two formulations are kept on purpose:
  * ``*_str`` functions work on Python strings exactly like the reference
    (slow, for small cases and for pinning the integer formulation);
  * integer-code functions (LUT → base-|A| codes) scale to the synthetic
    benchmark configs and are what the CUDA path is compared with.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

INVALID = 0xFF

# ---------------------------------------------------------------------------
# Alphabets — restates snekmer/alphabet.py:31-85 (ALPHABETS) as
# name -> [(input residues, output symbol), ...] in the reference's dict order
# (order matters: later groups overwrite earlier ones, alphabet.py:89-96).
# ---------------------------------------------------------------------------
_AA20 = "AILMVFYWSTQNCHDEKRGP"            # alphabet.py:13
_PTM = "-_!^#$@.%&"                        # alphabet.py:17
_GROUPS: Dict[str, List[Tuple[str, str]]] = {
    "hydro": [("SFTNKYEQCWPHDR", "S"), ("VMLAIG", "V")],
    "standard": [("AGILMV", "A"), ("PH", "P"), ("FWY", "F"), ("NQST", "N"),
                 ("DE", "D"), ("KR", "K"), ("C", "C")],
    "solvacc": [("CILMVFWY", "C"), ("AGHST", "A"), ("PDEKNQR", "P")],
    # E is absent and N is listed twice (second wins) in the reference.
    "hydrocharge": [("SFTNYQCWPH", "L"), ("VMLAIG", "H"), ("KNDR", "C")],
    "hydrostruct": [("SFTNKYEQCWHDR", "L"), ("VMLAI", "H"), ("PG", "B")],
    "miqs": [("A", "A"), ("C", "C"), ("DEN", "D"), ("FWY", "F"), ("G", "G"),
             ("H", "H"), ("ILMQV", "I"), ("KR", "K"), ("P", "P"), ("ST", "S")],
    "ptm": [(c, c) for c in _AA20 + _PTM],
    "None": [(c, c) for c in _AA20],
}
_ORDER = {0: "hydro", 1: "standard", 2: "solvacc", 3: "hydrocharge",
          4: "hydrostruct", 5: "miqs"}        # alphabet.py:21-28


def alphabet_name(alphabet) -> str:
    """alphabet.py:121-155 (check_valid) + :158-197 (get_alphabet) name lookup."""
    if alphabet is None:
        return "None"
    if isinstance(alphabet, (int, np.integer)) and not isinstance(alphabet, bool):
        if int(alphabet) in _ORDER:
            return _ORDER[int(alphabet)]
        if int(alphabet) in range(len(_GROUPS)):   # 6, 7 pass check_valid then KeyError
            raise KeyError(int(alphabet))
        raise ValueError("Invalid alphabet specified")
    if alphabet in _GROUPS:
        return alphabet
    raise ValueError("Invalid alphabet specified")


def residue_map(alphabet, extra: Optional[Dict[str, List[Tuple[str, str]]]] = None) -> Dict[str, str]:
    """Long-form residue→symbol dict, alphabet.py:88-96 (FULL_ALPHABETS)."""
    groups = extra[alphabet] if (extra and alphabet in extra) else _GROUPS[alphabet_name(alphabet)]
    m: Dict[str, str] = {}
    for src, dst in groups:
        for ch in src:
            m[ch] = dst
    return m


def symbols_of(alphabet, extra=None) -> str:
    """Output symbol set (alphabet.py:241-266 get_alphabet_keys), in the
    canonical *sorted* order used for integer codes."""
    return "".join(sorted(set(residue_map(alphabet, extra).values())))


def build_lut(alphabet, extra=None) -> Tuple[np.ndarray, str]:
    """256-entry byte LUT: residue byte → symbol index, 0xFF = invalid.

    Follows vectorize.py:193-195 (translate leaves unmapped characters
    unchanged) and :247 (a window is kept iff all its characters are in the
    output symbol set): a byte is valid iff its *translated* character is an
    output symbol, which also covers a raw unmapped byte that happens to equal
    an output symbol (raw 'B' under hydrostruct)."""
    m = residue_map(alphabet, extra)
    syms = symbols_of(alphabet, extra)
    lut = np.full(256, INVALID, dtype=np.uint8)
    for b in range(128):
        ch = chr(b)
        t = m.get(ch, ch)
        i = syms.find(t)
        if i >= 0 and len(t) == 1:
            lut[b] = i
    return lut, syms


# ---------------------------------------------------------------------------
# String-level restatement (small cases)
# ---------------------------------------------------------------------------
def reduce_str(sequence: str, alphabet, extra=None) -> str:
    """vectorize.py:173-195: rstrip('*') then one-pass translate."""
    s = str(sequence).rstrip("*")
    m = residue_map(alphabet, extra)
    return "".join(m.get(c, c) for c in s)


def reduce_vectorize_str(sequence: str, alphabet, k: int, extra=None) -> List[str]:
    """vectorize.py:292-328 + :239-249: valid k-mers in order, with repeats."""
    red = reduce_str(sequence, alphabet, extra)
    cs = set(symbols_of(alphabet, extra))
    return [red[i:i + k] for i in range(len(red) - k + 1) if set(red[i:i + k]) <= cs]


def basis_str(seqs: Iterable[str], alphabet, k: int, min_filter: int = 0, extra=None) -> List[str]:
    """kmerize.smk:89-104: first-occurrence order, keep total count > min_filter."""
    counts: Dict[str, int] = {}
    for s in seqs:
        for km in reduce_vectorize_str(s, alphabet, k, extra):
            counts[km] = counts.get(km, 0) + 1
    return [km for km, c in counts.items() if c > min_filter]


def counts_str(seq: str, kmerlist: Sequence[str], alphabet, extra=None) -> List[int]:
    """learn.smk:359-383 / apply.smk:195-206: all windows of the reduced string
    (no validity test), gathered in basis order."""
    if len(kmerlist) == 0:
        return []
    red = reduce_str(seq, alphabet, extra)
    k = len(kmerlist[0])
    d: Dict[str, int] = {}
    for i in range(len(red) - k + 1):
        w = red[i:i + k]
        d[w] = d.get(w, 0) + 1
    return [d.get(km, 0) for km in kmerlist]


# ---------------------------------------------------------------------------
# Integer-code formulation
# ---------------------------------------------------------------------------
def pack(seqs: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate sequences into one uint8 buffer + int64[N+1] offsets."""
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).copy()
    return buf, offs


def code_dtype(nsym: int, k: int):
    return np.uint32 if nsym ** k <= 2 ** 32 else np.uint64


def window_codes(residues: np.ndarray, offsets: np.ndarray, lut: np.ndarray, nsym: int, k: int):
    """All windows that lie inside one sequence.

    Returns (seq_idx, pos, code uint64, valid) per window start, in file order
    (sequence-major, position-minor) — the order reduce_vectorize emits k-mers
    (vectorize.py:239-249).  code = sum sym_i * nsym**(k-1-i); valid iff every
    symbol != 0xFF."""
    if nsym ** k > 2 ** 64:
        raise ValueError("|A|^k does not fit uint64")
    n = len(offsets) - 1
    lens = np.diff(offsets)
    nwin = np.maximum(lens - k + 1, 0)
    total = int(nwin.sum())
    seq_idx = np.repeat(np.arange(n, dtype=np.int64), nwin)
    wstart = np.zeros(n + 1, dtype=np.int64)
    wstart[1:] = np.cumsum(nwin)
    pos = np.arange(total, dtype=np.int64) - wstart[seq_idx]
    gpos = offsets[seq_idx] + pos
    sym = lut[residues]
    code = np.zeros(total, dtype=np.uint64)
    valid = np.ones(total, dtype=bool)
    base = np.uint64(nsym)
    for i in range(k):
        s = sym[gpos + i]
        valid &= s != INVALID
        code = code * base + np.where(s == INVALID, 0, s).astype(np.uint64)
    return seq_idx, pos, code, valid


def decode(codes: np.ndarray, symbols: str, k: int) -> np.ndarray:
    """Integer codes → k-mer strings ('<Uk')."""
    codes = np.asarray(codes, dtype=np.uint64)
    nsym = np.uint64(len(symbols))
    chars = np.empty((len(codes), k), dtype="<U1")
    sym = np.array(list(symbols))
    c = codes.copy()
    for i in range(k - 1, -1, -1):
        chars[:, i] = sym[(c % nsym).astype(np.int64)]
        c //= nsym
    if len(codes) == 0:
        return np.array([], dtype=f"<U{k}")
    return np.array(["".join(r) for r in chars], dtype=f"<U{k}")


def encode_kmers(kmers: Sequence[str], symbols: str) -> np.ndarray:
    """k-mer strings → integer codes (uint64); raises on foreign symbols."""
    idx = {c: i for i, c in enumerate(symbols)}
    out = np.zeros(len(kmers), dtype=np.uint64)
    n = len(symbols)
    for j, km in enumerate(kmers):
        v = 0
        for ch in km:
            v = v * n + idx[ch]
        out[j] = v
    return out


def basis_codes(seq_idx, pos, code, valid, min_filter: int = 0):
    """kmerize.smk:89-104 on codes: distinct valid codes, total count, ordered
    by first (seq, pos); keep count > min_filter.  Returns (codes, counts)."""
    c = code[valid]
    if c.size == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.int64)
    uniq, first, cnt = np.unique(c, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    uniq, cnt = uniq[order], cnt[order]
    keep = cnt > min_filter
    return uniq[keep], cnt[keep].astype(np.int64)


def count_matrix(seq_idx, code, valid, nseq: int, basis: np.ndarray) -> np.ndarray:
    """Dense int32 [nseq, K] counts in basis-column order
    (learn.smk:359-383; windows whose code is not in the basis are dropped)."""
    K = len(basis)
    out = np.zeros((nseq, K), dtype=np.int32)
    if K == 0:
        return out
    order = np.argsort(basis, kind="stable")
    sb = basis[order]
    c = code[valid]
    s = seq_idx[valid]
    j = np.searchsorted(sb, c)
    j[j >= K] = K - 1
    hit = sb[j] == c
    col = order[j[hit]]
    np.add.at(out, (s[hit], col), 1)
    return out


def count_csr(seq_idx, code, valid, nseq: int):
    """Per-sequence sorted distinct valid codes + counts (CSR in code space)."""
    c = code[valid]
    s = seq_idx[valid]
    order = np.lexsort((c, s))
    c, s = c[order], s[order]
    if c.size == 0:
        return np.zeros(nseq + 1, np.int64), np.zeros(0, np.uint64), np.zeros(0, np.int32)
    new = np.ones(c.size, dtype=bool)
    new[1:] = (c[1:] != c[:-1]) | (s[1:] != s[:-1])
    starts = np.flatnonzero(new)
    cnt = np.diff(np.append(starts, c.size)).astype(np.int32)
    rows = s[starts]
    rowptr = np.zeros(nseq + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, c[starts], cnt


def presence_matrix(counts: np.ndarray) -> np.ndarray:
    """kmerize.smk:112-120: float64 0/1."""
    return (counts > 0).astype(np.float64)


_ACC = re.compile(r"\|(.*?)\|")


def accession(seq_id: str) -> str:
    """learn.smk:322 — text between the first two pipes; IndexError otherwise."""
    return _ACC.findall(seq_id)[0]


def learn_matrix(ids: Sequence[str], counts: np.ndarray, seq_annot: Dict[str, str]):
    """learn.smk:306-357 (Library) on a dense per-sequence count matrix.

    Returns (annotations in first-appearance order, M int64 [A,K],
    seq_count int64 [A], totals int64 [K], total_seqs).
    Duplicate sequence ids: the reference keys a dict by id, so the last
    duplicate's counts win for the annotation rows while *all* duplicates have
    already been added to Totals (learn.smk:380, :312)."""
    K = counts.shape[1]
    totals = counts.astype(np.int64).sum(axis=0)
    last: Dict[str, int] = {}
    for i, sid in enumerate(ids):
        last[sid] = i                     # dict insertion order = first appearance
    ann_rows: Dict[str, np.ndarray] = {}
    ann_nseq: Dict[str, int] = {}
    for sid, i in last.items():
        acc = accession(sid)
        if acc not in seq_annot:
            continue
        a = seq_annot[acc]
        if a not in ann_rows:
            ann_rows[a] = np.zeros(K, dtype=np.int64)
            ann_nseq[a] = 0
        ann_rows[a] += counts[i]
        ann_nseq[a] += 1
    anns = list(ann_rows.keys())
    M = np.stack([ann_rows[a] for a in anns]) if anns else np.zeros((0, K), np.int64)
    nseq = np.array([ann_nseq[a] for a in anns], dtype=np.int64)
    return anns, M, nseq, totals, len(last)


def cosine_scores(Q: np.ndarray, M: np.ndarray, q_norm_sq: Optional[np.ndarray] = None) -> np.ndarray:
    """apply.smk:278-289 / learn.smk:811-829 via sklearn's definition:
    normalise rows (zero-norm rows stay zero), then Qn @ Mn.T, float64.
    ``q_norm_sq`` lets the caller supply ||q||^2 over a *wider* column set (the
    union re-index of apply.smk:268-276 keeps query k-mers the learned matrix
    does not have)."""
    Q = np.asarray(Q, dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)
    qn = np.sqrt((Q * Q).sum(axis=1) if q_norm_sq is None else np.asarray(q_norm_sq, np.float64))
    mn = np.sqrt((M * M).sum(axis=1))
    qn[qn == 0] = 1.0
    mn[mn == 0] = 1.0
    return (Q / qn[:, None]) @ (M / mn[:, None]).T


def top2(S: np.ndarray):
    """apply.smk:312-313 ``argsort(-S)[:, :2]`` with ties → lowest column.
    Returns (idx1, idx2, s1, s2); with a single column idx2 = -1, s2 = nan."""
    S = np.asarray(S)
    q, a = S.shape
    order = np.argsort(-S, axis=1, kind="stable")[:, :2]
    i1 = order[:, 0]
    s1 = S[np.arange(q), i1]
    if a >= 2:
        i2 = order[:, 1]
        s2 = S[np.arange(q), i2]
    else:
        i2 = np.full(q, -1)
        s2 = np.full(q, np.nan)
    return i1, i2, s1, s2


def apply_table(S: np.ndarray, annotations: Sequence[str], confidence: Optional[Dict[float, float]] = None):
    """apply.smk:301-340 restated (the original indexing at :317-319 raises
    under pandas 3): Prediction, Score, delta=round(top1-top2, 2), Confidence."""
    i1, i2, s1, s2 = top2(S)
    delta = np.round(s1 - s2, 2)
    pred = [str(annotations[i]) for i in i1]
    conf = None
    if confidence is not None:
        conf = np.array([confidence.get(float(d), np.nan) for d in delta])
    return pred, s1, delta, conf


# ---------------------------------------------------------------------------
# Convenience: the whole path on packed input, integer formulation
# ---------------------------------------------------------------------------
def vectorize_packed(residues, offsets, alphabet, k, min_filter=0, basis=None, extra=None):
    """(basis codes, total counts, dense count matrix) for a packed input."""
    lut, syms = build_lut(alphabet, extra)
    si, pos, code, valid = window_codes(residues, offsets, lut, len(syms), k)
    if basis is None:
        basis, tot = basis_codes(si, pos, code, valid, min_filter)
    else:
        tot = None
    C = count_matrix(si, code, valid, len(offsets) - 1, np.asarray(basis, np.uint64))
    return np.asarray(basis, np.uint64), tot, C
