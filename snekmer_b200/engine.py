"""engine: batch entry points of the B200 hot path (host side of the C-ABI).

PyTorch is used for device memory, streams and (elsewhere) torch.distributed
only; every computation below is a call into libskm_b200.so with raw device
pointers.  All functions raise if CUDA or the library is unavailable — there is
no CPU fallback.

Data layout in HBM
  residues  uint8 [R]    all sequences back to back (ASCII), 16-byte aligned
  offsets   int64 [N+1]  sequence s is residues[offsets[s]:offsets[s+1]]
  codes     uint32/uint64 base-|A| k-mer codes (symbols in sorted order)
  basis     uint64 [K] codes in first-occurrence order (kmerize.smk:89-104)
  col_of_code int32 [|A|^k]  code → basis column, -1 if filtered out
  counts    int32/uint16 [N, K] row-major, or CSR (rowptr int64, cols uint32, vals int32)
  M         int64 [A+1, K] annotation rows + one "rest" row; totals int64 [K]
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native
from . import alphabet as _alphabet
from ._native import SkmError, check, lib

INT64_MAX = (1 << 63) - 1

AlphabetT = Union[str, int, None]


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise SkmError(-101, "no CUDA device: snekmer_b200 has no CPU path (the oracle lives in oracle/ and is test-only)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise SkmError(-101, f"device {dev} is not a CUDA device")
    return dev


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class AlphabetTables:
    name: str
    symbols: str
    nsym: int
    lut_host: bytes
    lut: torch.Tensor          # uint8 [256] on device
    charmap: torch.Tensor      # uint8 [256] on device


_tables_cache = {}


def alphabet_tables(alphabet, device=None) -> AlphabetTables:
    """Device tables of an alphabet given by name / index / None — or an AlphabetTables
    object, which is passed through (lets callers use ad-hoc symbol sets)."""
    if isinstance(alphabet, AlphabetTables):
        return alphabet
    dev = _require_cuda(device)
    name = _alphabet.get_alphabet_name(alphabet)
    key = (name, dev.index, tuple(sorted(_alphabet.residue_map(name).items())))
    hit = _tables_cache.get(key)
    if hit is not None:
        return hit
    syms = _alphabet.symbols(name)
    lut_host = _alphabet.lut(name)
    t = AlphabetTables(
        name=name, symbols=syms, nsym=len(syms), lut_host=lut_host,
        lut=torch.frombuffer(bytearray(lut_host), dtype=torch.uint8).to(dev),
        charmap=torch.frombuffer(bytearray(_alphabet.charmap(name)), dtype=torch.uint8).to(dev),
    )
    _tables_cache[key] = t
    return t


def alphabet_tables_from_symbols(symbols: str, device=None) -> AlphabetTables:
    """Identity alphabet over already REDUCED text: every character of `symbols` is its own
    symbol, everything else is invalid.  This is what the learn / apply rules need: they
    count windows of the reduced strings stored in the .npz against a k-mer list
    (learn.smk:359-383, apply.smk:188-206), so a window can only match when all its
    characters occur in that list."""
    dev = _require_cuda(device)
    syms = "".join(sorted(set(symbols)))
    key = ("=" + syms, dev.index)
    hit = _tables_cache.get(key)
    if hit is not None:
        return hit
    lut_host = _native.lut_build(syms, syms, syms)
    t = AlphabetTables(name="=" + syms, symbols=syms, nsym=len(syms), lut_host=lut_host,
                       lut=torch.frombuffer(bytearray(lut_host), dtype=torch.uint8).to(dev),
                       charmap=torch.arange(256, dtype=torch.uint8, device=dev))
    _tables_cache[key] = t
    return t


def code_space(nsym: int, k: int) -> int:
    return nsym ** k


class SequenceBatch:
    """Packed sequences resident in HBM."""

    def __init__(self, residues: torch.Tensor, offsets: torch.Tensor, offsets_host: np.ndarray):
        assert residues.dtype == torch.uint8 and offsets.dtype == torch.int64
        self.residues = residues
        self.offsets = offsets
        self.offsets_host = offsets_host
        self.n = len(offsets_host) - 1
        self.nres = int(offsets_host[-1]) if self.n >= 0 and len(offsets_host) else 0
        lens = np.diff(offsets_host) if self.n > 0 else np.zeros(0, np.int64)
        self.max_len = int(lens.max()) if lens.size else 0

    @property
    def device(self):
        return self.residues.device

    @staticmethod
    def pack_host(seqs: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        offs = np.zeros(len(seqs) + 1, dtype=np.int64)
        if len(seqs):
            np.cumsum([len(s) for s in seqs], out=offs[1:])
        buf = np.frombuffer("".join(seqs).encode("latin-1", "replace"), dtype=np.uint8)
        return buf, offs

    @classmethod
    def from_packed(cls, residues: np.ndarray, offsets: np.ndarray, device=None, pinned: bool = False) -> "SequenceBatch":
        dev = _require_cuda(device)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        nres = int(offsets[-1]) if len(offsets) else 0
        # +16 bytes of slack keeps the buffer 16-byte granular; kernels never read past nres
        d_res = torch.empty(max(nres, 1) + 16, dtype=torch.uint8, device=dev)
        if nres:
            src = torch.from_numpy(residues[:nres] if residues.flags.writeable else residues[:nres].copy())
            d_res[:nres].copy_(src, non_blocking=pinned)
        d_off = torch.from_numpy(offsets).to(dev)
        return cls(d_res, d_off, offsets)

    @classmethod
    def from_strings(cls, seqs: Sequence[str], device=None) -> "SequenceBatch":
        buf, offs = cls.pack_host([str(s) for s in seqs])
        return cls.from_packed(buf, offs, device)


# ---------------------------------------------------------------------------
# (a) encode
# ---------------------------------------------------------------------------
def encode_windows(batch: SequenceBatch, alphabet: AlphabetT, k: int) -> torch.Tensor:
    """Code of the window starting at every residue (KmerVec.reduce_vectorize,
    vectorize.py:292-328, for the whole batch).  Returns uint32-as-int32 or
    uint64-as-int64 tensor [nres]; invalid windows are all-ones (-1)."""
    tab = alphabet_tables(alphabet, batch.device)
    S = code_space(tab.nsym, k)
    bits = 32 if S < 2 ** 32 else 64
    out = torch.empty(max(batch.nres, 1), dtype=torch.int32 if bits == 32 else torch.int64, device=batch.device)
    check(lib().skm_encode_windows(_ptr(batch.residues), batch.nres, _ptr(batch.offsets), batch.n, _ptr(tab.lut),
                                   tab.nsym, int(k), bits, _ptr(out), _stream()))
    return out[:batch.nres]


def reduce_bytes(batch: SequenceBatch, alphabet: AlphabetT) -> torch.Tensor:
    """reduce() (vectorize.py:173-195) on the packed buffer: translated bytes [nres]."""
    tab = alphabet_tables(alphabet, batch.device)
    out = torch.empty_like(batch.residues)
    check(lib().skm_reduce_bytes(_ptr(batch.residues), batch.nres, _ptr(tab.charmap), _ptr(out), _stream()))
    return out[:batch.nres]


# ---------------------------------------------------------------------------
# (b') basis
# ---------------------------------------------------------------------------
@dataclass
class Basis:
    """k-mer basis in first-occurrence order + the code→column map."""
    alphabet: str
    k: int
    symbols: str
    codes: torch.Tensor                 # int64 (uint64 bit pattern) [K]
    counts: Optional[torch.Tensor]      # int64 [K] total occurrences, None for a supplied basis
    col_of_code: torch.Tensor           # int32 [S]
    K: int
    S: int

    def codes_host(self) -> np.ndarray:
        return self.codes.cpu().numpy().view(np.uint64)

    def kmers(self) -> np.ndarray:
        return decode_kmers(self.codes_host(), self.symbols, self.k)


def decode_kmers(codes: np.ndarray, symbols: str, k: int) -> np.ndarray:
    """uint64 codes → '<Uk' strings (host-side formatting of the basis)."""
    codes = np.asarray(codes, dtype=np.uint64)
    if codes.size == 0:
        return np.array([], dtype=f"<U{max(k, 1)}")
    n = np.uint64(len(symbols))
    sym = np.frombuffer(symbols.encode("latin-1"), dtype=np.uint8)
    chars = np.empty((codes.size, k), dtype=np.uint8)
    c = codes.copy()
    for i in range(k - 1, -1, -1):
        chars[:, i] = sym[(c % n).astype(np.int64)]
        c //= n
    return chars.view(f"S{k}").ravel().astype(f"<U{k}")


def encode_kmers(kmers: Sequence[str], symbols: str, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """k-mer strings → (uint64 codes, ok mask).  ok is False for k-mers of the
    wrong length or with foreign symbols (they can never match a window)."""
    arr = np.asarray(list(kmers), dtype=str)
    ok = np.ones(arr.size, dtype=bool)
    codes = np.zeros(arr.size, dtype=np.uint64)
    if arr.size == 0:
        return codes, ok
    idx = np.full(256, -1, dtype=np.int64)
    for i, ch in enumerate(symbols):
        idx[ord(ch)] = i
    lens = np.char.str_len(arr)
    ok &= lens == k
    n = np.uint64(len(symbols))
    safe = np.where(ok, arr, symbols[0] * k)
    try:
        raw = np.char.encode(safe, "latin-1")
    except UnicodeEncodeError:
        raw = np.array([s.encode("latin-1", "replace") for s in safe])
    mat = np.frombuffer(raw.astype(f"S{k}").tobytes(), dtype=np.uint8).reshape(arr.size, k)
    d = idx[mat]
    ok &= (d >= 0).all(axis=1)
    d = np.where(d < 0, 0, d).astype(np.uint64)
    for i in range(k):
        codes = codes * n + d[:, i]
    return codes, ok


def basis_tables(S: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fresh (count, first) accumulation tables over the code space."""
    count = torch.zeros(S, dtype=torch.int64, device=device)
    first = torch.full((S,), -1, dtype=torch.int64, device=device)   # all-ones = +inf as uint64
    return count, first


def basis_accumulate(batch: SequenceBatch, alphabet: AlphabetT, k: int, count: torch.Tensor, first: Optional[torch.Tensor],
                     res_base: int = 0) -> None:
    """count[c] += occurrences of code c; first[c] = min(first[c], global position) unless first is None
    (occurrence counts only: the Totals row of learn.smk:380)."""
    tab = alphabet_tables(alphabet, batch.device)
    check(lib().skm_basis_accumulate(_ptr(batch.residues), batch.nres, _ptr(batch.offsets), batch.n, _ptr(tab.lut),
                                     tab.nsym, int(k), int(res_base), _ptr(count), _ptr(first), _stream()))


def kmer_totals(batch: SequenceBatch, alphabet: AlphabetT, k: int) -> torch.Tensor:
    """int64 [nsym^k]: occurrences of every k-mer code over ALL sequences of the batch (learn.smk:380 Totals)."""
    tab = alphabet_tables(alphabet, batch.device)
    S = code_space(tab.nsym, k)
    if S > _native.SKM_DENSE_MAX_SPACE:
        raise SkmError(-3, f"code space {tab.nsym}^{k} exceeds the table limit 2^27")
    count = torch.zeros(S, dtype=torch.int64, device=batch.device)
    basis_accumulate(batch, alphabet, k, count, None, 0)
    return count


def basis_finalize(alphabet: AlphabetT, k: int, count: Optional[torch.Tensor], first: torch.Tensor, min_filter: int = 0) -> Basis:
    """Tables -> basis in first-occurrence order.  count may be None (order-only tables, min_filter 0): every code
    with a first position is kept and Basis.counts is None."""
    dev = first.device
    tab = alphabet_tables(alphabet, dev)
    S = first.numel()
    ws_bytes = lib().skm_basis_finalize_workspace(S)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    codes = torch.empty(S, dtype=torch.int64, device=dev)
    counts = torch.empty(S, dtype=torch.int64, device=dev) if count is not None else None
    col = torch.empty(S, dtype=torch.int32, device=dev)
    dK = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().skm_basis_finalize(_ptr(count), _ptr(first), S, int(min_filter), _ptr(codes), _ptr(counts), _ptr(col),
                                   _ptr(dK), _ptr(ws), ws_bytes, _stream()))
    K = int(dK.item())
    return Basis(tab.name, int(k), tab.symbols, codes[:K], None if counts is None else counts[:K], col, K, S)


ORDER_FIRST_CHUNK_RES = 4 << 20     # residues in the first chunk of the order-only basis walk
ORDER_GROWTH = 4                    # every further chunk is this many times longer


def basis_first_progressive(batch: SequenceBatch, alphabet: AlphabetT, k: int, first: torch.Tensor, state: torch.Tensor,
                            res_base: int = 0) -> None:
    """first[c] = min(first[c], global position of the first window with code c), walking the shard front to back
    and stopping on the device once every code of the space has been seen (state int32[4], zeroed by the caller)."""
    tab = alphabet_tables(alphabet, batch.device)
    oh = np.ascontiguousarray(batch.offsets_host, dtype=np.int64)
    check(lib().skm_basis_first_progressive(_ptr(batch.residues), batch.nres, _ptr(batch.offsets), oh.ctypes.data, batch.n,
                                            _ptr(tab.lut), tab.nsym, int(k), int(res_base), _ptr(first), _ptr(state),
                                            ORDER_FIRST_CHUNK_RES, ORDER_GROWTH, _stream()))


def order_only_supported(S: int) -> bool:
    return S <= lib().skm_basis_order_max_space()


def build_basis(batch: SequenceBatch, alphabet: AlphabetT, k: int, min_filter: int = 0, counts: bool = True) -> Basis:
    """Pass 1 of the vectorize rule (kmerize.smk:89-104) for one shard.

    counts=False with min_filter <= 0 asks for the basis ORDER only (Basis.counts is None): the reference uses the
    occurrence counts for nothing but the `count > min_filter` test, and with min_filter = 0 the basis is determined
    by the first positions alone, so the walk stops as soon as every code of the space has been seen (small code
    spaces; others take the full pass)."""
    tab = alphabet_tables(alphabet, batch.device)
    S = code_space(tab.nsym, k)
    if S > _native.SKM_DENSE_MAX_SPACE:
        raise SkmError(-3, f"code space {tab.nsym}^{k} exceeds the table limit 2^27")
    if not counts and min_filter <= 0 and order_only_supported(S):
        first = torch.full((S,), -1, dtype=torch.int64, device=batch.device)
        state = torch.zeros(4, dtype=torch.int32, device=batch.device)
        basis_first_progressive(batch, alphabet, k, first, state, 0)
        return basis_finalize(alphabet, k, None, first, 0)
    count, first = basis_tables(S, batch.device)
    basis_accumulate(batch, alphabet, k, count, first, 0)
    return basis_finalize(alphabet, k, count, first, min_filter)


class OrderOnlyPlan:
    """The launch-bound front of the vectorize rule for ONE resident batch — table reset, order-only basis walk
    (4 chunk launches), key build, sort and emit of the finalisation: ten tiny kernels, ~0.15 ms of host-paced launches in
    front of a 0.85 ms count pass — captured once into a CUDA graph and replayed as a single launch.  The graph holds
    the batch's pointers and chunk geometry, so a plan belongs to its batch (a resident upload buffer that is refilled
    with batches of the same layout keeps its plan only if the offsets are the same; otherwise build a new one or run
    without).  Capture failure (a driver that refuses a node) leaves `graph = None`: the calls then run eagerly."""

    def __init__(self, batch: SequenceBatch, alphabet: AlphabetT, k: int, capture: bool = True):
        tab = alphabet_tables(alphabet, batch.device)
        dev = batch.device
        S = code_space(tab.nsym, k)
        if not order_only_supported(S):
            raise SkmError(-3, f"OrderOnlyPlan: code space {tab.nsym}^{k} exceeds {lib().skm_basis_order_max_space()}")
        self.batch, self.alphabet, self.k, self.tab, self.S = batch, alphabet, int(k), tab, S
        self.first = torch.empty(S, dtype=torch.int64, device=dev)
        self.state = torch.empty(4, dtype=torch.int32, device=dev)
        self.ws_bytes = lib().skm_basis_finalize_workspace(S)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.codes = torch.empty(S, dtype=torch.int64, device=dev)
        self.col = torch.empty(S, dtype=torch.int32, device=dev)
        self.dK = torch.empty(1, dtype=torch.int64, device=dev)
        self.graph = None
        if capture:
            self.enqueue()                               # eager once: function attributes, lazy module loading
            torch.cuda.synchronize(dev)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.enqueue()
                self.graph = g
            except Exception:                            # noqa: BLE001 — stay on the eager path
                self.graph = None
                torch.cuda.synchronize(dev)

    def enqueue(self) -> None:
        self.first.fill_(-1)
        self.state.zero_()
        self.dK.zero_()
        basis_first_progressive(self.batch, self.alphabet, self.k, self.first, self.state, 0)
        check(lib().skm_basis_finalize(None, _ptr(self.first), self.S, 0, _ptr(self.codes), None, _ptr(self.col), _ptr(self.dK),
                                       _ptr(self.ws), self.ws_bytes, _stream()))

    def launch(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self.enqueue()


def vectorize_order_only(batch: SequenceBatch, alphabet: AlphabetT, k: int, out: Optional[torch.Tensor] = None,
                         dtype: torch.dtype = torch.int32, count_events=None, plan: Optional[OrderOnlyPlan] = None) -> Tuple[Basis, torch.Tensor]:
    """Both passes of the vectorize rule (kmerize.smk:89-120) for min_filter = 0 over a small code space
    (order_only_supported): order-only basis walk + dense counts, with the read-back of K taken OFF the critical path.
    K only sizes the output, and K <= S: the count pass is launched with S columns (the column map sends every basis
    k-mer below K, the columns from K on stay zero) before the host asks for K, so the 8-byte read-back and the host
    latency around it overlap the count kernel instead of leaving the GPU idle between the passes (~0.07 ms of a
    1.04 ms C2 step).  Returns (basis, counts [N, K]); when the space is not saturated (K < S) the counts are compacted
    to K columns.  `out`: optional [N, S] buffer; `count_events`: optional (start, end) CUDA events recorded around the
    count launch; `plan`: an OrderOnlyPlan of THIS batch (the front of the step as one CUDA-graph launch).  The basis
    tensors of the result alias the plan's buffers: they are valid until the plan is launched again."""
    tab = alphabet_tables(alphabet, batch.device)
    dev = batch.device
    S = code_space(tab.nsym, k)
    if plan is None:
        plan = OrderOnlyPlan(batch, alphabet, k, capture=False)
    else:
        assert plan.batch is batch and plan.S == S and plan.k == int(k) and plan.tab.name == tab.name
    plan.launch()
    codes, col, dK = plan.codes, plan.col, plan.dK
    bits = {torch.int32: 32, torch.uint16: 16, torch.int16: 16}[dtype]
    if out is None:
        out = torch.empty((batch.n, S), dtype=dtype, device=dev)
    else:
        assert out.is_contiguous() and tuple(out.shape) == (batch.n, S) and out.dtype == dtype
    if count_events:
        count_events[0].record()
    check(lib().skm_count_dense(_ptr(batch.residues), batch.nres, _ptr(batch.offsets), batch.n, _ptr(tab.lut), tab.nsym,
                                int(k), _ptr(col), S, S, bits, _ptr(out), batch.max_len, _stream()))
    if count_events:
        count_events[1].record()
    K = int(dK.item())                                  # the count pass is already running
    basis = Basis(tab.name, int(k), tab.symbols, codes[:K], None, col, K, S)
    if K < S:
        out = out[:, :K].contiguous()
    return basis, out


def basis_from_kmers(kmers: Sequence[str], alphabet: AlphabetT, k: int, device=None) -> Tuple[Basis, np.ndarray]:
    """A supplied basis (basis.txt branch, kmerize.smk:72-78, or a learned
    kmerlist).  Returns (Basis over the encodable k-mers, index of each of them
    in the supplied list)."""
    dev = _require_cuda(device)
    tab = alphabet_tables(alphabet, dev)
    S = code_space(tab.nsym, k)
    if S > _native.SKM_DENSE_MAX_SPACE:
        raise SkmError(-3, f"code space {tab.nsym}^{k} exceeds the table limit 2^27")
    codes, ok = encode_kmers(kmers, tab.symbols, k)
    keep = np.flatnonzero(ok)
    d_codes = torch.from_numpy(codes[keep].view(np.int64)).to(dev)
    col = torch.empty(S, dtype=torch.int32, device=dev)
    check(lib().skm_basis_colmap(_ptr(d_codes), len(keep), S, _ptr(col), _stream()))
    return Basis(tab.name, int(k), tab.symbols, d_codes, None, col, len(keep), S), keep


def gather_columns(matrix, index) -> "np.ndarray | torch.Tensor":
    """out[:, j] = matrix[:, index[j]], zero where index[j] is outside [0, n) — the column
    gather of KmerBasis.transform (vectorize.py:54-119).  numpy in → numpy out, tensor in → tensor out."""
    as_numpy = not isinstance(matrix, torch.Tensor)
    dev = _require_cuda(None if as_numpy else matrix.device)
    if as_numpy:
        arr = np.ascontiguousarray(matrix)
        if arr.ndim == 1:
            arr = arr.reshape(1, -1)
        if arr.dtype.itemsize not in (1, 2, 4, 8) or arr.dtype.kind not in "iufb":
            raise TypeError(f"gather_columns: unsupported dtype {arr.dtype}")
        d_in = torch.from_numpy(arr.view(f"u{arr.dtype.itemsize}").view(np.dtype(f"i{arr.dtype.itemsize}"))).to(dev)
    else:
        d_in = matrix.contiguous()
    rows, n = d_in.shape
    idx = torch.as_tensor(np.asarray(index, dtype=np.int64) if not isinstance(index, torch.Tensor) else index,
                          dtype=torch.int64, device=dev).contiguous()
    out = torch.empty((rows, idx.numel()), dtype=d_in.dtype, device=dev)
    check(lib().skm_gather_columns(_ptr(d_in), rows, n, d_in.element_size(), _ptr(idx), idx.numel(), _ptr(out), _stream()))
    if as_numpy:
        return out.cpu().numpy().view(arr.dtype)
    return out


def scatter_add(dst: torch.Tensor, src: torch.Tensor, row_map, col_map) -> None:
    """dst[row_map[r], col_map[c]] += src[r, c] for int64 matrices (negative map entries are dropped):
    the outer-join-and-sum of Merge.merge_dataframes (learn.smk:467-494)."""
    dev = dst.device
    assert dst.dtype == torch.int64 and src.dtype == torch.int64 and dst.is_contiguous()
    src = src.contiguous()
    rm = torch.as_tensor(row_map, dtype=torch.int64, device=dev).contiguous()
    cm = torch.as_tensor(col_map, dtype=torch.int64, device=dev).contiguous()
    assert rm.numel() == src.shape[0] and cm.numel() == src.shape[1]
    check(lib().skm_scatter_add_i64(_ptr(src), src.shape[0], src.shape[1], _ptr(rm), _ptr(cm), _ptr(dst), dst.shape[0],
                                    dst.shape[1], _stream()))


def basis_from_codes(codes: np.ndarray, alphabet, k: int, device=None) -> Tuple[Basis, np.ndarray]:
    """Basis whose column j holds code codes[j] (uint64, distinct)."""
    dev = _require_cuda(device)
    tab = alphabet_tables(alphabet, dev)
    S = code_space(tab.nsym, k)
    if S > _native.SKM_DENSE_MAX_SPACE:
        raise SkmError(-3, f"code space {tab.nsym}^{k} exceeds the table limit 2^27")
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    d_codes = torch.from_numpy(codes.view(np.int64)).to(dev)
    col = torch.empty(S, dtype=torch.int32, device=dev)
    check(lib().skm_basis_colmap(_ptr(d_codes), len(codes), S, _ptr(col), _stream()))
    return Basis(tab.name, int(k), tab.symbols, d_codes, None, col, len(codes), S), np.arange(len(codes))


def intersect_basis(target: Basis, other: Basis) -> Basis:
    """Columns of `target`, restricted to codes that `other` also holds.

    apply.smk:268-276 re-indexes query counts (columns = the query file's own
    basis) onto the union of columns: a learned column receives a query count
    only if the k-mer is in both bases."""
    assert target.S == other.S
    minus1 = torch.full_like(target.col_of_code, -1)
    col = torch.where(other.col_of_code >= 0, target.col_of_code, minus1)
    return Basis(target.alphabet, target.k, target.symbols, target.codes, target.counts, col, target.K, target.S)


# ---------------------------------------------------------------------------
# (b) counts
# ---------------------------------------------------------------------------
def count_dense(batch: SequenceBatch, alphabet: AlphabetT, k: int, basis: Optional[Basis] = None,
                dtype: torch.dtype = torch.int32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-sequence k-mer counts [N, K] over `basis` (None = identity basis: column = code)."""
    tab = alphabet_tables(alphabet, batch.device)
    S = code_space(tab.nsym, k)
    K = S if basis is None else basis.K
    bits = {torch.int32: 32, torch.uint16: 16, torch.int16: 16}[dtype]
    if out is None:
        out = torch.empty((batch.n, K), dtype=dtype, device=batch.device)
    else:
        assert out.is_contiguous() and tuple(out.shape) == (batch.n, K) and out.dtype == dtype
    check(lib().skm_count_dense(_ptr(batch.residues), batch.nres, _ptr(batch.offsets), batch.n, _ptr(tab.lut), tab.nsym,
                                int(k), None if basis is None else _ptr(basis.col_of_code), S, K, bits, _ptr(out),
                                batch.max_len, _stream()))
    return out


def count_over_kmers(batch: SequenceBatch, alphabet, k: int, kmerlist: Sequence[str], dtype: torch.dtype = torch.int32) -> torch.Tensor:
    """Counts [N, len(kmerlist)] with one column per LIST ENTRY, in list order
    (``[k_counts.get(kmer, 0) for kmer in kmerlist]``, learn.smk:377-382): entries that
    cannot be encoded (wrong length, foreign characters) are zero columns, repeated
    entries repeat the column."""
    basis, keep = basis_from_kmers(kmerlist, alphabet, k, batch.device)
    n_list = len(kmerlist)
    codes = basis.codes_host()
    if basis.K == n_list and np.unique(codes).size == n_list:
        return count_dense(batch, alphabet, k, basis, dtype=dtype)
    # distinct codes -> dense columns, then expand to the list
    uniq, inverse = np.unique(codes, return_inverse=True)
    ub, _ = basis_from_codes(uniq, alphabet, k, batch.device)
    C = count_dense(batch, alphabet, k, ub, dtype=dtype)
    index = np.full(n_list, -1, dtype=np.int64)
    index[keep] = inverse
    return gather_columns(C, index)


def count_csr(batch: SequenceBatch, alphabet: AlphabetT, k: int, basis: Optional[Basis] = None, method: str = "warp"):
    """Per-sequence sorted (column, count) pairs as CSR: (rowptr int64 [N+1], cols int32 [nnz], vals int32 [nnz]).
    method "warp" (default): a warp sorts one sequence's keys in shared memory (skm_count_csr_sorted);
    "segsort": window keys in HBM + cub::DeviceSegmentedSort (skm_count_csr, the first implementation, kept as a
    cross-check)."""
    tab = alphabet_tables(alphabet, batch.device)
    S = code_space(tab.nsym, k)
    dev = batch.device
    rowptr = torch.empty(batch.n + 1, dtype=torch.int64, device=dev)
    cols = torch.empty(max(batch.nres, 1), dtype=torch.int32, device=dev)
    vals = torch.empty(max(batch.nres, 1), dtype=torch.int32, device=dev)
    if method == "segsort":
        ws_bytes = lib().skm_count_csr_workspace(batch.nres, batch.n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib().skm_count_csr(_ptr(batch.residues), batch.nres, _ptr(batch.offsets), batch.n, _ptr(tab.lut), tab.nsym,
                                  int(k), None if basis is None else _ptr(basis.col_of_code), S, _ptr(rowptr), _ptr(cols),
                                  _ptr(vals), _ptr(ws), ws_bytes, _stream()))
    else:
        ws_bytes = lib().skm_count_csr_sorted_workspace(batch.nres, batch.n, batch.max_len, 32)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib().skm_count_csr_sorted(_ptr(batch.residues), batch.nres, _ptr(batch.offsets), batch.n, _ptr(tab.lut), tab.nsym,
                                         int(k), 32, None if basis is None else _ptr(basis.col_of_code), S, None, None, 0,
                                         batch.max_len, _ptr(rowptr), _ptr(cols), None, _ptr(vals), _ptr(ws), ws_bytes, _stream()))
    nnz = int(rowptr[-1].item()) if batch.n else 0
    return rowptr, cols[:nnz].clone(), vals[:nnz].clone()


# ---------------------------------------------------------------------------
# (c) learn
# ---------------------------------------------------------------------------
def learn_dense(batch: SequenceBatch, alphabet: AlphabetT, k: int, basis: Optional[Basis], ann_id: torch.Tensor,
                n_ann: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-annotation summed counts.  ann_id int32 [N] (<0 = not in any row).
    Returns (M int64 [n_ann+1, K] with the "rest" row last, totals int64 [K])."""
    tab = alphabet_tables(alphabet, batch.device)
    S = code_space(tab.nsym, k)
    K = S if basis is None else basis.K
    dev = batch.device
    ann_id = ann_id.to(device=dev, dtype=torch.int32).contiguous()
    assert ann_id.numel() == batch.n
    # group equal ids: stable sort so that each annotation is one contiguous run
    order = torch.sort(torch.where(ann_id < 0, torch.full_like(ann_id, n_ann), ann_id), stable=True).indices.contiguous()
    M = torch.empty((n_ann + 1, K), dtype=torch.int64, device=dev)
    totals = torch.empty(K, dtype=torch.int64, device=dev)
    check(lib().skm_learn_dense(_ptr(batch.residues), batch.nres, _ptr(batch.offsets), batch.n, _ptr(tab.lut), tab.nsym,
                                int(k), None if basis is None else _ptr(basis.col_of_code), S, K, _ptr(ann_id),
                                _ptr(order), int(n_ann), _ptr(M), _ptr(totals), _stream()))
    return M, totals


def learn_over_kmers(batch: SequenceBatch, alphabet, k: int, kmerlist: Sequence[str], ann_id: torch.Tensor,
                     n_ann: int) -> Tuple[np.ndarray, np.ndarray]:
    """learn_dense over a k-mer LIST: (M int64 [n_ann, len(kmerlist)], totals int64 [len(kmerlist)]) on the
    host, one column per list entry (zero columns for entries no window can match)."""
    basis, keep = basis_from_kmers(kmerlist, alphabet, k, batch.device)
    codes = basis.codes_host()
    n_list = len(kmerlist)
    plain = basis.K == n_list and np.unique(codes).size == n_list
    if not plain:
        uniq, inverse = np.unique(codes, return_inverse=True)
        basis, _ = basis_from_codes(uniq, alphabet, k, batch.device)
    M, totals = learn_dense(batch, alphabet, k, basis, ann_id, n_ann)
    if not plain:
        index = np.full(n_list, -1, dtype=np.int64)
        index[keep] = inverse
        M = gather_columns(M, index)
        totals = gather_columns(totals.reshape(1, -1), index).reshape(-1)
    return M[:n_ann].cpu().numpy(), totals.cpu().numpy()


# ---------------------------------------------------------------------------
# (d) apply
# ---------------------------------------------------------------------------
def row_norm2(X: torch.Tensor) -> torch.Tensor:
    X = X.contiguous()
    out = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
    fn = {torch.int32: lib().skm_row_norm2_i32, torch.int64: lib().skm_row_norm2_i64}[X.dtype]
    check(fn(_ptr(X), X.shape[0], X.shape[1], _ptr(out), _stream()))
    return out


@dataclass
class ApplyResult:
    top1: torch.Tensor      # int32 [Q]
    top2: torch.Tensor      # int32 [Q] (-1 if fewer than 2 annotations)
    score1: torch.Tensor    # float64 [Q]
    score2: torch.Tensor    # float64 [Q]
    scores: Optional[torch.Tensor] = None   # float64 [Q, A] when requested


TC_MAX_K = 32768


@dataclass
class PreparedAnnotations:
    """An annotation matrix split into base-256 digit planes for the tensor-core scoring kernel."""
    planes: torch.Tensor        # uint8 buffer [n_planes, pad128(A), pad128(K)]
    n_planes: int
    n_ann: int
    K: int
    mnorm2: torch.Tensor        # float64 [A]
    M: Optional[torch.Tensor] = None    # the int64 matrix itself (a reference, not a copy): per-row fallback of apply_tc

    def tiles(self) -> np.ndarray:
        """The annotation tiles of the prepared matrix as int32 [n_tiles, 3] = (first sorted row, width, digit planes)
        (blob header: n_tiles, rows per plane; TileDesc records follow the row permutation)."""
        hdr = self.planes[:8].view(torch.int32).cpu().numpy()
        n_tiles, rows = int(hdr[0]), int(hdr[1])
        off = 1024 + rows * 4
        return self.planes[off:off + 16 * n_tiles].view(torch.int32).cpu().numpy().reshape(-1, 4)[:, :3].copy()

    def issued_macs_per_query(self) -> int:
        """int8 multiply-accumulates apply_tc issues per query row: sum over tiles of width x planes x padded K."""
        t = self.tiles()
        return int((t[:, 1].astype(np.int64) * t[:, 2]).sum()) * ((self.K + 127) // 128 * 128)


def prepare_annotations(M: torch.Tensor, mnorm2: Optional[torch.Tensor] = None) -> Optional[PreparedAnnotations]:
    """Digit planes of M for apply_tc, or None when M is outside the tensor-core envelope
    (K > 32768, entries >= 2^32 or negative).  Do this once per learned matrix."""
    M = M.contiguous()
    A, K = M.shape
    if A == 0 or K == 0 or K > TC_MAX_K:
        return None
    dev = _require_cuda(M.device)
    nbytes = lib().skm_apply_tc_planes_bytes(A, K)
    planes = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    n_planes = _native.C.c_int(0)
    rc = lib().skm_apply_tc_prepare(_ptr(M), A, K, _ptr(planes), nbytes, _native.C.byref(n_planes), _stream())
    if rc == -3:            # SKM_ERR_UNSUPPORTED: too many digit planes
        return None
    check(rc)
    return PreparedAnnotations(planes, int(n_planes.value), A, K, row_norm2(M) if mnorm2 is None else mnorm2, M)


def rows_out_of_range(X: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """Sorted indices (int64) of the rows of an int32 matrix that hold an element outside [lo, hi]."""
    dev = _require_cuda(X.device)
    X = X.contiguous()
    rows, cols = X.shape
    out = torch.empty(max(rows, 1), dtype=torch.int32, device=dev)
    dn = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().skm_rows_out_of_range_i32(_ptr(X), rows, cols, int(lo), int(hi), _ptr(out), rows, _ptr(dn), _stream()))
    n = int(dn.item())
    return torch.sort(out[:n].to(torch.int64)).values


def apply_tc(Q: torch.Tensor, prep: PreparedAnnotations, qnorm2: Optional[torch.Tensor] = None,
             full: bool = False) -> Optional[ApplyResult]:
    """Tensor-core scoring (tcgen05 int8 GEMM, exact integer dots).  Query rows holding a count above 255 (a
    5,000-residue low-complexity protein) do not fit the uint8 operand: they — and only they — are re-scored by the
    exact CUDA-core path (skm_apply_dense) against prep.M.  Returns None only when such rows exist and the
    prepared matrix does not carry M."""
    dev = _require_cuda(Q.device)
    Q = Q.contiguous()
    nq, K = Q.shape
    assert K == prep.K and Q.dtype == torch.int32
    if qnorm2 is None:
        qnorm2 = row_norm2(Q)
    A = prep.n_ann
    top1 = torch.empty(nq, dtype=torch.int32, device=dev)
    top2 = torch.empty(nq, dtype=torch.int32, device=dev)
    s1 = torch.empty(nq, dtype=torch.float64, device=dev)
    s2 = torch.empty(nq, dtype=torch.float64, device=dev)
    scores = torch.empty((nq, A), dtype=torch.float64, device=dev) if full else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib().skm_apply_tc_workspace(nq, K, A)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib().skm_apply_tc(_ptr(Q), nq, K, _ptr(prep.planes), prep.n_planes, A, _ptr(qnorm2), _ptr(prep.mnorm2), _ptr(top1),
                             _ptr(top2), _ptr(s1), _ptr(s2), _ptr(scores), _ptr(status), _ptr(ws), ws_bytes, _stream()))
    if int(status.item()) != 0:
        if prep.M is None:
            return None
        bad = rows_out_of_range(Q, 0, 255)
        r = apply_dense(Q[bad].contiguous(), prep.M, qnorm2=qnorm2[bad].contiguous(), mnorm2=prep.mnorm2, full=full,
                        tensor_cores=False)
        top1[bad], top2[bad], s1[bad], s2[bad] = r.top1, r.top2, r.score1, r.score2
        if full:
            scores[bad] = r.scores
    return ApplyResult(top1, top2, s1, s2, scores)


def apply_dense(Q: torch.Tensor, M: torch.Tensor, qnorm2: Optional[torch.Tensor] = None,
                mnorm2: Optional[torch.Tensor] = None, full: bool = False, chunk: int = 1 << 16,
                tensor_cores: Optional[bool] = None, prepared: Optional[PreparedAnnotations] = None) -> ApplyResult:
    """Cosine of every query count row against every annotation row + top-2
    (apply.smk:278-335).  Q int32 [nq, K], M int64 [A, K]; qnorm2 defaults to the
    norm over Q's own columns.  tensor_cores: None = use the tcgen05 path when the
    operands fit its exactness envelope and the problem is large enough to pay for the
    operand split; True / False force the choice (True still falls back when the envelope
    is violated)."""
    dev = Q.device
    Q = Q.contiguous()
    M = M.contiguous()
    nq, K = Q.shape
    A = M.shape[0]
    assert M.shape[1] == K and Q.dtype == torch.int32 and M.dtype == torch.int64
    if qnorm2 is None:
        qnorm2 = row_norm2(Q)
    if mnorm2 is None:
        mnorm2 = row_norm2(M) if prepared is None else prepared.mnorm2
    want_tc = tensor_cores if tensor_cores is not None else (nq * A * K >= (1 << 24))
    if want_tc and nq > 0 and A > 0 and 0 < K <= TC_MAX_K:
        prep = prepared if prepared is not None else prepare_annotations(M, mnorm2)
        if prep is not None:
            r = apply_tc(Q, prep, qnorm2, full)
            if r is not None:
                return r
    top1 = torch.empty(nq, dtype=torch.int32, device=dev)
    top2 = torch.empty(nq, dtype=torch.int32, device=dev)
    s1 = torch.empty(nq, dtype=torch.float64, device=dev)
    s2 = torch.empty(nq, dtype=torch.float64, device=dev)
    scores = torch.empty((nq, A), dtype=torch.float64, device=dev) if full else None
    chunk = max(1, min(chunk, nq, (1 << 31) // max(1, A * 8)))      # dots workspace <= 2 GiB
    ws_bytes = lib().skm_apply_dense_workspace(chunk, A, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    for q0 in range(0, nq, chunk):
        q1 = min(nq, q0 + chunk)
        check(lib().skm_apply_dense(_ptr(Q[q0:q1]), q1 - q0, K, _ptr(M), A, _ptr(qnorm2[q0:q1]), _ptr(mnorm2),
                                    _ptr(top1[q0:q1]), _ptr(top2[q0:q1]), _ptr(s1[q0:q1]), _ptr(s2[q0:q1]),
                                    None if scores is None else _ptr(scores[q0:q1]), _ptr(ws), ws_bytes, _stream()))
    return ApplyResult(top1, top2, s1, s2, scores)


def merge_top2(idx: torch.Tensor, score: torch.Tensor) -> ApplyResult:
    """Global top-2 from stacked per-shard top-2 lists (idx int64 [W, 2, Q] global annotation
    indices or -1, score float64 [W, 2, Q]) — the fan-in of annotation-sharded apply."""
    dev = _require_cuda(idx.device)
    idx = idx.contiguous()
    score = score.contiguous()
    W, two, nq = idx.shape
    assert two == 2 and tuple(score.shape) == (W, 2, nq) and idx.dtype == torch.int64 and score.dtype == torch.float64
    top1 = torch.empty(nq, dtype=torch.int32, device=dev)
    top2 = torch.empty(nq, dtype=torch.int32, device=dev)
    s1 = torch.empty(nq, dtype=torch.float64, device=dev)
    s2 = torch.empty(nq, dtype=torch.float64, device=dev)
    check(lib().skm_top2_merge(_ptr(idx), _ptr(score), W, nq, _ptr(top1), _ptr(top2), _ptr(s1), _ptr(s2), _stream()))
    return ApplyResult(top1, top2, s1, s2, None)


# ---------------------------------------------------------------------------
# sparse paths (large bases): sort-based learn, CSC, SpMM apply
# ---------------------------------------------------------------------------
def _sub_batch(batch: SequenceBatch, lo: int, hi: int) -> SequenceBatch:
    """Sequences [lo, hi) of a batch as a self-contained shard (same HBM buffer, rebased offsets)."""
    oh = batch.offsets_host
    base = int(oh[lo]) & ~15                      # keep the 16-byte alignment of the residue pointer
    sub = SequenceBatch(batch.residues[base:], batch.offsets[lo:hi + 1] - base, oh[lo:hi + 1] - base)
    sub.nres = int(oh[hi]) - base
    return sub


def _chunks_by_residues(offsets_host: np.ndarray, max_res: int):
    n = len(offsets_host) - 1
    out, lo = [], 0
    while lo < n:
        hi = int(np.searchsorted(offsets_host, offsets_host[lo] + max_res, side="right")) - 1
        hi = min(max(hi, lo + 1), n)
        out.append((lo, hi))
        lo = hi
    return out


def coo_merge(keys: torch.Tensor, vals: torch.Tensor, key_bound: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sum entries with equal keys; result sorted by key (Merge.merge_dataframes for sparse matrices).
    key_bound: exclusive upper bound of the keys (n_ann * S) when known — the sort then runs over its bits only."""
    dev = _require_cuda(keys.device)
    n = keys.numel()
    keys, vals = keys.contiguous(), vals.contiguous()
    ws_bytes = lib().skm_coo_merge_workspace(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ok = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    ov = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    dn = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().skm_coo_merge(_ptr(keys), _ptr(vals), n, int(key_bound), _ptr(ok), _ptr(ov), _ptr(dn), _ptr(ws), ws_bytes, _stream()))
    m = int(dn.item())
    return ok[:m].clone(), ov[:m].clone()


def gather_sequences(batch: SequenceBatch, sel: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sequences sel[0], sel[1], ... of a batch packed back to back: (residues uint8 [R' + 16], offsets int64 [n_sel + 1])."""
    dev = batch.device
    sel = sel.to(device=dev, dtype=torch.int64).contiguous()
    lens = (batch.offsets[1:] - batch.offsets[:-1])[sel]
    out_off = torch.zeros(sel.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=out_off[1:])
    total = int(out_off[-1].item()) if sel.numel() else 0
    out_res = torch.empty(total + 32, dtype=torch.uint8, device=dev)
    check(lib().skm_gather_sequences(_ptr(batch.residues), _ptr(batch.offsets), _ptr(sel), sel.numel(), _ptr(out_res), _ptr(out_off), _stream()))
    return out_res, out_off


def coo_merge_runs(keys: torch.Tensor, vals: torch.Tensor, run_sizes: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """coo_merge for an input that is a concatenation of SORTED runs (run_sizes entries each): merge tree + reduce-by-key."""
    dev = _require_cuda(keys.device)
    n = keys.numel()
    assert sum(run_sizes) == n
    keys, vals = keys.contiguous(), vals.contiguous()
    offs = np.zeros(len(run_sizes) + 1, dtype=np.int64)
    np.cumsum(np.asarray(run_sizes, dtype=np.int64), out=offs[1:])
    ws_bytes = lib().skm_coo_merge_runs_workspace(n, len(run_sizes))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ok = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    ov = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    dn = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().skm_coo_merge_runs(_ptr(keys), _ptr(vals), offs.ctypes.data, len(run_sizes), _ptr(ok), _ptr(ov), _ptr(dn), _ptr(ws),
                                   ws_bytes, _stream()))
    m = int(dn.item())
    return ok[:m].clone(), ov[:m].clone()


def coo_merge_runs_packed(packed_ptr: int, run_sizes: Sequence[int], count_bits: int, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """coo_merge_runs for SORTED runs of packed words (key << count_bits | count, skm_coo_pack's format) lying back to
    back at device address `packed_ptr` (a peer-memory receive buffer).  Returns (keys, vals, n) with n a device scalar:
    the first n entries are the merged list — the caller slices after its own synchronisation."""
    dev = _require_cuda(device)
    n = int(sum(run_sizes))
    offs = np.zeros(len(run_sizes) + 1, dtype=np.int64)
    np.cumsum(np.asarray(run_sizes, dtype=np.int64), out=offs[1:])
    ws_bytes = lib().skm_coo_merge_runs_packed_workspace(n, len(run_sizes))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ok = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    ov = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    dn = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().skm_coo_merge_runs_packed(packed_ptr, offs.ctypes.data, len(run_sizes), int(count_bits), _ptr(ok), _ptr(ov), _ptr(dn),
                                          _ptr(ws), ws_bytes, _stream()))
    return ok, ov, dn


def coo_pack(keys: torch.Tensor, vals: torch.Tensor, count_bits: int) -> Tuple[torch.Tensor, bool]:
    """(packed words int64, overflow) — skm_coo_pack; the single-GPU half of the exchange format (tests, side-cars)."""
    dev = _require_cuda(keys.device)
    out = torch.empty(keys.numel(), dtype=torch.int64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().skm_coo_pack(_ptr(keys.contiguous()), _ptr(vals.contiguous()), keys.numel(), int(count_bits), _ptr(out), _ptr(flag), _stream()))
    return out, bool(flag.item())


def learn_sparse(batch: SequenceBatch, alphabet, k: int, ann_id: torch.Tensor, n_ann: int,
                 max_chunk_res: int = 1 << 28, method: str = "grouped", place: Optional[dict] = None):
    """Annotation x k-mer count matrix as a COO list sorted by key = ann * S + code
    (keys int64 holding the uint64 pattern, vals int64).  Sequences with ann_id < 0 do not
    contribute (Totals come from the basis tables).

    method "grouped" (default): the annotated sequences are gathered in annotation order and sorted one annotation
    slice at a time with 32-bit keys (skm_learn_sparse_group); "global": one 64-bit radix sort over all window keys of
    the shard (skm_learn_sparse, the first implementation, kept as a cross-check and for slices whose keys need more
    than 32 bits).
    place (grouped path only; learn_sparse_hybrid): dict(ins_ann, ins_cum, ins_pos int64 [H], totals int64 [S] or None,
    extra int) — the list is written straight into a larger list with room for H foreign blocks
    (skm_learn_sparse_group_place); returns (keys buffer, vals buffer, entries of THIS list) without slicing, or None when
    the grouped path does not apply (the caller then places the list itself)."""
    tab = alphabet_tables(alphabet, batch.device)
    dev = batch.device
    ann_id = ann_id.to(device=dev, dtype=torch.int32).contiguous()
    assert ann_id.numel() == batch.n
    S = code_space(tab.nsym, k)
    per_group = ((1 << 32) - 2) // S
    if method == "grouped" and per_group >= 1 and batch.n > 0 and n_ann > 0:
        key = torch.where((ann_id >= 0) & (ann_id < n_ann), ann_id, torch.full_like(ann_id, n_ann))
        sk, order = torch.sort(key, stable=True)
        starts = list(range(0, n_ann, per_group)) + [n_ann]
        bounds = torch.searchsorted(sk, torch.tensor(starts, dtype=torch.int32, device=dev)).tolist()     # one sync
        n_sel = bounds[-1]
        if n_sel == 0:
            if place:
                cap = max(int(place["extra"]), 1)
                return torch.empty(cap, dtype=torch.int64, device=dev), torch.empty(cap, dtype=torch.int64, device=dev), 0
            z = torch.zeros(0, dtype=torch.int64, device=dev)
            return z, z.clone()
        g_res, g_off = gather_sequences(batch, order[:n_sel])
        g_ann = sk[:n_sel].contiguous()
        edge = g_off[torch.tensor(bounds, dtype=torch.int64, device=dev)].tolist()
        if max(edge[i + 1] - (edge[i] & ~15) for i in range(len(starts) - 1)) < (1 << 31) - 64:
            total = edge[-1] + (int(place["extra"]) if place else 0)
            keys = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
            vals = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
            dn = torch.zeros(1, dtype=torch.int64, device=dev)
            ws = None
            for gi in range(len(starts) - 1):
                s0, s1 = bounds[gi], bounds[gi + 1]
                if s1 == s0:
                    continue
                base = edge[gi] & ~15
                nres = edge[gi + 1] - base
                ws_bytes = lib().skm_learn_sparse_group_workspace(nres)
                if ws is None or ws.numel() < ws_bytes:
                    ws = None
                    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                offs = (g_off[s0:s1 + 1] - base).contiguous()
                if place:
                    check(lib().skm_learn_sparse_group_place(_ptr(g_res[base:]), nres, _ptr(offs), s1 - s0, _ptr(tab.lut), tab.nsym, int(k),
                                                             _ptr(g_ann[s0:s1]), starts[gi], starts[gi + 1] - starts[gi], _ptr(keys), _ptr(vals),
                                                             total, _ptr(dn), _ptr(place["ins_ann"]), _ptr(place["ins_cum"]),
                                                             int(place["ins_ann"].numel()), _ptr(place["ins_pos"]), _ptr(place.get("totals")),
                                                             _ptr(ws), ws.numel(), _stream()))
                else:
                    check(lib().skm_learn_sparse_group(_ptr(g_res[base:]), nres, _ptr(offs), s1 - s0, _ptr(tab.lut), tab.nsym, int(k),
                                                       _ptr(g_ann[s0:s1]), starts[gi], starts[gi + 1] - starts[gi], _ptr(keys), _ptr(vals),
                                                       total, _ptr(dn), _ptr(ws), ws.numel(), _stream()))
            m = int(dn.item())
            assert m <= total
            if place:
                return keys, vals, m
            return keys[:m].clone(), vals[:m].clone()
    if place:
        return None
    parts_k, parts_v = [], []
    for lo, hi in _chunks_by_residues(batch.offsets_host, max_chunk_res):
        sub = _sub_batch(batch, lo, hi)
        ws_bytes = lib().skm_learn_sparse_workspace(sub.nres)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ok = torch.empty(max(sub.nres, 1), dtype=torch.int64, device=dev)
        ov = torch.empty(max(sub.nres, 1), dtype=torch.int64, device=dev)
        dn = torch.zeros(1, dtype=torch.int64, device=dev)
        check(lib().skm_learn_sparse(_ptr(sub.residues), sub.nres, _ptr(sub.offsets), sub.n, _ptr(tab.lut), tab.nsym, int(k),
                                     _ptr(ann_id[lo:hi]), int(n_ann), _ptr(ok), _ptr(ov), _ptr(dn), _ptr(ws), ws_bytes,
                                     _stream()))
        m = int(dn.item())
        parts_k.append(ok[:m].clone())
        parts_v.append(ov[:m].clone())
        del ws, ok, ov
    if len(parts_k) == 1:
        return parts_k[0], parts_v[0]
    if not parts_k:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, z.clone()
    return coo_merge(torch.cat(parts_k), torch.cat(parts_v), key_bound=n_ann * S)


HEAVY_ROW_DIV = int(os.environ.get("SKM_HEAVY_ROW_DIV", "4"))   # an annotation is "heavy" when it has more residues than S / HEAVY_ROW_DIV ...
HEAVY_ROWS_MAX_BYTES = 8 << 30   # ... as long as the dense rows of all heavy annotations stay below this
ROWS_L2_BYTES = int(os.environ.get("SKM_ROWS_L2_MB", "100")) << 20   # rows counted by one skm_rows_accumulate launch (L2-resident counters): ~0.8 of the 126 MB L2, from a sweep (profiles/R2al_*)


def _learn_rows(batch: SequenceBatch, tab: AlphabetTables, k: int, ann_id: torch.Tensor, n_ann: int, S: int,
                want_totals: bool, heavy_div: int, min_res: int = 0, heavy_ids: Optional[torch.Tensor] = None):
    """Dense-row part of learn_sparse_hybrid: returns (rows uint32 [R, S], heavy annotation ids int64 [H] ascending,
    rest_row index or -1).  Rows 0..H-1 are the heavy annotations, row H (when want_totals) the unannotated sequences."""
    dev = batch.device
    lens = (batch.offsets[1:] - batch.offsets[:-1])
    valid = (ann_id >= 0) & (ann_id < n_ann)
    per_ann = torch.zeros(n_ann + 1, dtype=torch.int64, device=dev)
    per_ann.index_add_(0, torch.where(valid, ann_id, torch.full_like(ann_id, n_ann)).to(torch.int64), lens)
    if heavy_ids is not None:                            # the caller chose the rows (ascending ids)
        heavy = heavy_ids.to(device=dev, dtype=torch.int64)
    else:
        thr = max(S // max(int(heavy_div), 1), 1, int(min_res))
        heavy_mask = per_ann[:n_ann] > thr
        max_rows = max(int(HEAVY_ROWS_MAX_BYTES // (4 * S)) - 1, 0)
        heavy = torch.nonzero(heavy_mask).reshape(-1)
        if heavy.numel() > max_rows:                         # keep the largest ones
            order = torch.argsort(per_ann[heavy], descending=True)[:max_rows]
            heavy = torch.sort(heavy[order]).values
    H = int(heavy.numel())                               # one sync
    rest_row = H if want_totals else -1
    R = H + (1 if want_totals else 0)
    if R == 0:
        return None, heavy, -1
    row_of_ann = torch.full((n_ann + 1,), -1, dtype=torch.int32, device=dev)
    row_of_ann[heavy] = torch.arange(H, dtype=torch.int32, device=dev)
    row_of_ann[n_ann] = rest_row
    row_of_seq = row_of_ann[torch.where(valid, ann_id, torch.full_like(ann_id, n_ann)).to(torch.int64)].contiguous()
    # the sequences of a row next to each other: the row stays in L2 while they stream past
    sel = torch.nonzero(row_of_seq >= 0).reshape(-1)
    order = torch.sort(row_of_seq[sel].to(torch.int64), stable=True).indices
    sel = sel[order]
    rows = torch.zeros((R, S), dtype=torch.int32, device=dev)
    if sel.numel():
        g_res, g_off = gather_sequences(batch, sel)
        g_row = row_of_seq[sel].contiguous()
        total = int(g_off[-1].item())
        # One launch per group of rows whose counters fit L2 together (ROWS_L2_BYTES): a launch spreads its CTAs over the
        # whole residue range it is given, so a single launch over all rows touches every row at once (C3: 66 rows =
        # 442 MB) and the atomics run at DRAM speed.  Also < 2^32 residues per call.
        per_launch = max(int(ROWS_L2_BYTES // (4 * S)), 1)
        row_edges = torch.searchsorted(g_row, torch.arange(R + 1, dtype=torch.int32, device=dev)).tolist()
        g_off_host = None
        bounds = [0]
        for r0 in range(0, R, per_launch):
            lo, hi = row_edges[r0], row_edges[min(r0 + per_launch, R)]
            if hi == lo:
                continue
            if total >= (1 << 32) - 64:                  # a huge group: cut it by sequences
                if g_off_host is None:
                    g_off_host = g_off.cpu().numpy()
                while g_off_host[hi] - g_off_host[bounds[-1]] >= (1 << 31):
                    bounds.append(int(np.searchsorted(g_off_host, g_off_host[bounds[-1]] + (1 << 31), side="right")) - 1)
            bounds.append(hi)
        edges = g_off[torch.tensor(bounds, dtype=torch.int64, device=dev)].tolist()
        for i, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
            if hi == lo:
                continue
            e0, e1 = edges[i], edges[i + 1]
            base = e0 & ~15
            offs = (g_off[lo:hi + 1] - base).contiguous()
            check(lib().skm_rows_accumulate(_ptr(g_res[base:]), e1 - base, _ptr(offs), hi - lo, _ptr(tab.lut), tab.nsym, int(k),
                                            _ptr(g_row[lo:hi]), int(S), _ptr(rows), _stream()))
    return rows, heavy, rest_row


def learn_sparse_hybrid(batch: SequenceBatch, alphabet, k: int, ann_id: torch.Tensor, n_ann: int, want_totals: bool = False,
                        heavy_div: int = HEAVY_ROW_DIV):
    """learn_sparse for heavy-tailed family sizes: annotations with more residues than S / heavy_div are COUNTED into
    dense L2-resident rows (skm_rows_accumulate; their sequences gathered together), the others are sorted
    (skm_learn_sparse_group); the two sorted lists are interleaved by annotation.  Same result as learn_sparse, bit for
    bit.  want_totals: also return the Totals row (int64 [S], occurrences over ALL sequences, learn.smk:380) — the
    unannotated sequences are then counted into a row of their own and the totals are column sums."""
    tab = alphabet_tables(alphabet, batch.device)
    dev = batch.device
    S = code_space(tab.nsym, k)
    ann_id = ann_id.to(device=dev, dtype=torch.int32).contiguous()
    assert ann_id.numel() == batch.n
    if S > _native.SKM_DENSE_MAX_SPACE:
        raise SkmError(-3, f"code space {tab.nsym}^{k} exceeds the table limit 2^27")
    rows, heavy, rest_row = _learn_rows(batch, tab, k, ann_id, n_ann, S, want_totals, heavy_div) if batch.n and n_ann else (None, torch.zeros(0, dtype=torch.int64, device=dev), -1)
    H = int(heavy.numel())
    totals = torch.zeros(S, dtype=torch.int64, device=dev) if want_totals else None
    # light part: the heavy annotations (and, with totals, nothing else) leave the sorted path
    if H:
        is_heavy = torch.zeros(n_ann + 1, dtype=torch.bool, device=dev)
        is_heavy[heavy] = True
        safe = torch.where((ann_id >= 0) & (ann_id < n_ann), ann_id, torch.full_like(ann_id, n_ann)).to(torch.int64)
        ann_light = torch.where(is_heavy[safe], torch.full_like(ann_id, -1), ann_id)
    else:
        ann_light = ann_id
    if H and os.environ.get("SKM_LEARN_PLACE", "1") != "0":
        # The heavy blocks' sizes first (the rows are counted already), then the light list is sorted slice by slice and
        # written STRAIGHT to its final places between the heavy blocks, adding its counts to the Totals on the way
        # (skm_learn_sparse_group_place): no separate column-sum pass over the light list and no shift copy of it.
        blk = lib().skm_rows_block()
        nblk = (S + blk - 1) // blk
        counts = torch.empty((H, nblk), dtype=torch.int32, device=dev)
        check(lib().skm_rows_block_counts(_ptr(rows), H, S, _ptr(counts), _stream()))
        c64 = counts.to(torch.int64)
        incl = torch.cumsum(c64, dim=1)
        blk_off = (incl - c64).contiguous()
        row_nnz = incl[:, -1].contiguous()
        cum = torch.cumsum(row_nnz, 0).contiguous()          # inclusive
        heavy_total = int(cum[-1].item())                    # one sync
        ins_pos = torch.full((H,), INT64_MAX, dtype=torch.int64, device=dev)
        placed = learn_sparse(batch, alphabet, k, ann_light, n_ann,
                              place=dict(ins_ann=heavy.contiguous(), ins_cum=cum, ins_pos=ins_pos, totals=totals, extra=heavy_total))
        if placed is not None:
            keys, vals, m_light = placed
            if want_totals:
                check(lib().skm_rows_colsum(_ptr(rows), rows.shape[0], S, _ptr(totals), _stream()))
            # light entries in front of block h = the smallest recorded index over the blocks from h on (none: all of them)
            tail = torch.cat([ins_pos, torch.tensor([m_light], dtype=torch.int64, device=dev)])
            before = torch.flip(torch.cummin(torch.flip(tail, [0]), 0).values, [0])[:H]
            row_dst = (before + cum - row_nnz).contiguous()
            total = m_light + heavy_total
            check(lib().skm_rows_emit(_ptr(rows), H, S, _ptr(blk_off), _ptr(row_dst), _ptr(heavy.contiguous()), _ptr(keys), _ptr(vals), total, _stream()))
            keys, vals = keys[:total], vals[:total]
            return (keys, vals, totals) if want_totals else (keys, vals)
    kl, vl = learn_sparse(batch, alphabet, k, ann_light, n_ann)
    if want_totals:
        check(lib().skm_coo_colsum(_ptr(kl), _ptr(vl), kl.numel(), S, _ptr(totals), _stream()))
        if rows is not None:
            check(lib().skm_rows_colsum(_ptr(rows), rows.shape[0], S, _ptr(totals), _stream()))
        else:                                           # no row machinery (empty batch / no annotations): count the rest directly
            rest = torch.nonzero((ann_id < 0) | (ann_id >= n_ann)).reshape(-1)
            if rest.numel():
                g_res, g_off = gather_sequences(batch, rest)
                basis_accumulate(SequenceBatch(g_res, g_off, g_off.cpu().numpy()), alphabet, k, totals, None, 0)
    if H == 0:
        return (kl, vl, totals) if want_totals else (kl, vl)
    # heavy part: non-zeros per block -> places -> emit
    blk = lib().skm_rows_block()
    nblk = (S + blk - 1) // blk
    counts = torch.empty((H, nblk), dtype=torch.int32, device=dev)
    check(lib().skm_rows_block_counts(_ptr(rows), H, S, _ptr(counts), _stream()))
    c64 = counts.to(torch.int64)
    incl = torch.cumsum(c64, dim=1)
    blk_off = (incl - c64).contiguous()                  # exclusive inside a row
    row_nnz = incl[:, -1].contiguous()
    # where the heavy blocks go: position in the light list (entries of smaller annotations) + heavy entries before
    ins_pos = torch.searchsorted(kl, heavy * S).contiguous()
    cum = torch.cumsum(row_nnz, 0).contiguous()          # inclusive
    row_dst = (ins_pos + cum - row_nnz).contiguous()
    total = int(kl.numel()) + int(cum[-1].item())        # one sync
    keys = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    vals = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    check(lib().skm_coo_shift_copy(_ptr(kl), _ptr(vl), kl.numel(), _ptr(ins_pos), _ptr(cum), H, _ptr(keys), _ptr(vals), _stream()))
    check(lib().skm_rows_emit(_ptr(rows), H, S, _ptr(blk_off), _ptr(row_dst), _ptr(heavy.contiguous()), _ptr(keys), _ptr(vals), total, _stream()))
    keys, vals = keys[:total], vals[:total]
    return (keys, vals, totals) if want_totals else (keys, vals)


ONCHIP_HEAVY_SCALE = float(os.environ.get("SKM_ONCHIP_HEAVY_SCALE", "0.75"))   # dense rows beyond scale * sqrt(task capacity * S) residues
ONCHIP_FILL = 0.8                                                      # target fill of a task cut from the code histogram


def _hist_bins(nsym: int, k: int, S: int) -> Tuple[int, int]:
    """(n_bins, bin_width) of the per-annotation code histogram: the leading j symbols of a code, nsym^j <= 8192."""
    j = 0
    while j < k and nsym ** (j + 1) <= 8192:
        j += 1
    return nsym ** j, nsym ** (k - j)


def learn_sparse_onchip(batch: SequenceBatch, alphabet, k: int, ann_id: torch.Tensor, n_ann: int, want_totals: bool = False,
                        heavy_scale: float = ONCHIP_HEAVY_SCALE):
    """learn_sparse without a device-wide sort.  Annotations with more than heavy_scale * sqrt(task capacity * S) residues
    (where a dense row of S counters — ~20 bytes of traffic per code — costs less than the n^2 / capacity window scans of
    the tasks the annotation would be cut into; C3: 158 k residues) are
    counted into dense L2-resident rows (skm_rows_*); every other annotation is sorted ON CHIP by skm_ann_sort: one CTA
    per task (annotation x code range, at most skm_ann_sort_cap() windows) scans, radix-sorts and run-length-encodes in
    shared memory and writes its entries at their final place of the one sorted COO list.  Annotations larger than a task
    are cut into code ranges from an exact histogram (skm_ann_hist).  Same result as learn_sparse, bit for bit; falls
    back to learn_sparse_hybrid when a histogram bin alone exceeds a task (degenerate low-complexity input).
    want_totals: also the Totals row (learn.smk:380)."""
    tab = alphabet_tables(alphabet, batch.device)
    dev = batch.device
    S = code_space(tab.nsym, k)
    ann_id = ann_id.to(device=dev, dtype=torch.int32).contiguous()
    assert ann_id.numel() == batch.n
    if S > _native.SKM_DENSE_MAX_SPACE:
        raise SkmError(-3, f"code space {tab.nsym}^{k} exceeds the table limit 2^27")
    if batch.n == 0 or n_ann == 0:
        return learn_sparse_hybrid(batch, alphabet, k, ann_id, n_ann, want_totals=want_totals)
    T = int(lib().skm_ann_sort_cap())
    thr = max(T, int(heavy_scale * (float(T) * float(S)) ** 0.5))
    max_rows = max(int(HEAVY_ROWS_MAX_BYTES // (4 * S)) - 1, 0)
    ok_id = (ann_id >= 0) & (ann_id < n_ann)
    safe = torch.where(ok_id, ann_id, torch.full_like(ann_id, n_ann)).to(torch.int64)
    lens = batch.offsets[1:] - batch.offsets[:-1]
    per_ann = torch.zeros(n_ann + 1, dtype=torch.int64, device=dev)
    per_ann.index_add_(0, safe, lens)
    heavy = torch.nonzero(per_ann[:n_ann] > thr).reshape(-1)
    if heavy.numel() > max_rows:                                      # keep the largest ones as rows
        heavy = torch.sort(heavy[torch.argsort(per_ann[heavy], descending=True)[:max_rows]]).values
    # ---- task annotations: sequences gathered in annotation order -------------------------------------------------------
    is_heavy = torch.zeros(n_ann + 1, dtype=torch.bool, device=dev)
    is_heavy[heavy] = True
    key = torch.where(ok_id & ~is_heavy[safe], ann_id, torch.full_like(ann_id, n_ann))
    sk, order = torch.sort(key, stable=True)
    uniq, per = torch.unique_consecutive(sk, return_counts=True)
    if uniq.numel() and int(uniq[-1].item()) == n_ann:                # the unselected sequences sort last
        uniq, per = uniq[:-1], per[:-1]
    A = int(uniq.numel())
    n_tasks = 0
    if A:
        seq_hi = torch.cumsum(per, 0)
        seq_lo = seq_hi - per
        n_sel = int(seq_hi[-1].item())
        g_res, g_off = gather_sequences(batch, order[:n_sel])
        nres_g = int(g_off[-1].item())
        ares = g_off[seq_hi] - g_off[seq_lo]                          # residues per task annotation
        medium = ares > T
        mi = torch.nonzero(medium).reshape(-1)
        M = int(mi.numel())
        t_ann = [uniq[~medium].to(torch.int64)]
        t_lo, t_hi = [seq_lo[~medium]], [seq_hi[~medium]]
        t_c0 = [torch.zeros(A - M, dtype=torch.int64, device=dev)]
        t_c1 = [torch.full((A - M,), S, dtype=torch.int64, device=dev)]
        if M:
            n_bins, width = _hist_bins(tab.nsym, k, S)
            hist = torch.empty((M, n_bins), dtype=torch.int32, device=dev)
            m_lo, m_hi = seq_lo[mi].to(torch.int32).contiguous(), seq_hi[mi].to(torch.int32).contiguous()
            check(lib().skm_ann_hist(_ptr(g_res), nres_g, _ptr(g_off), n_sel, _ptr(tab.lut), tab.nsym, int(k), _ptr(m_lo), _ptr(m_hi), M,
                                     width, n_bins, _ptr(hist), _stream()))
            fill = int(T * ONCHIP_FILL)
            # an annotation with one bin beyond what a task can absorb (low-complexity families) gets a dense row instead
            hot = hist.max(dim=1).values.to(torch.int64) > T - fill
            n_hot = int(hot.sum().item())
            if n_hot:
                heavy = torch.sort(torch.cat([heavy, uniq[mi][hot].to(torch.int64)])).values
                if heavy.numel() > max_rows:
                    return learn_sparse_hybrid(batch, alphabet, k, ann_id, n_ann, want_totals=want_totals)
                keep = torch.nonzero(~hot).reshape(-1)
                mi, hist = mi[keep], hist[keep]
                M = int(mi.numel())
        if M:
            h64 = hist.to(torch.int64)
            cum = torch.cumsum(h64, 1)
            q = torch.div((cum - 1).clamp_(min=0), fill, rounding_mode="floor")
            start = torch.ones_like(q, dtype=torch.bool)
            start[:, 1:] = q[:, 1:] != q[:, :-1]
            rb = torch.nonzero(start)                                 # (medium row, first bin) of every task, row-major
            r_, b_ = rb[:, 0], rb[:, 1]
            b_hi = torch.full_like(b_, n_bins)
            b_hi[:-1] = torch.where(r_[1:] == r_[:-1], b_[1:], b_hi[:-1])
            t_ann.append(uniq[mi][r_].to(torch.int64))
            t_lo.append(seq_lo[mi][r_])
            t_hi.append(seq_hi[mi][r_])
            t_c0.append(b_ * width)
            t_c1.append(torch.clamp(b_hi * width, max=S))
        t_ann, t_lo, t_hi, t_c0, t_c1 = (torch.cat(x) for x in (t_ann, t_lo, t_hi, t_c0, t_c1))
        perm = torch.argsort(t_ann * (S + 1) + t_c0)
        t_ann = t_ann[perm].contiguous()
        t_lo, t_hi = t_lo[perm].to(torch.int32).contiguous(), t_hi[perm].to(torch.int32).contiguous()
        t_c0, t_c1 = t_c0[perm].to(torch.int32).contiguous(), t_c1[perm].to(torch.int32).contiguous()     # uint32 patterns (S <= 2^27)
        n_tasks = int(t_ann.numel())
    # ---- heavy rows (+ the row of the unannotated sequences): counted, then entries per row -------------------------------
    rows, heavy, rest_row = _learn_rows(batch, tab, k, ann_id, n_ann, S, want_totals, heavy_div=S + 1, heavy_ids=heavy)
    H = int(heavy.numel())
    totals = torch.zeros(S, dtype=torch.int64, device=dev) if want_totals else None
    if H:
        blk = lib().skm_rows_block()
        nblk = (S + blk - 1) // blk
        counts = torch.empty((H, nblk), dtype=torch.int32, device=dev)
        check(lib().skm_rows_block_counts(_ptr(rows), H, S, _ptr(counts), _stream()))
        c64 = counts.to(torch.int64)
        incl = torch.cumsum(c64, dim=1)
        blk_off = (incl - c64).contiguous()
        row_nnz = incl[:, -1].contiguous()
        heavy_cum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(row_nnz, 0)])    # exclusive, H + 1
    else:
        heavy_cum = torch.zeros(1, dtype=torch.int64, device=dev)
    if n_tasks:
        out_base = heavy_cum[torch.searchsorted(heavy, t_ann)].contiguous() if H else torch.zeros(n_tasks, dtype=torch.int64, device=dev)
    heavy_total = int(heavy_cum[-1].item()) if H else 0
    capacity = heavy_total + (nres_g if A else 0)
    keys = torch.empty(max(capacity, 1), dtype=torch.int64, device=dev)
    vals = torch.empty(max(capacity, 1), dtype=torch.int64, device=dev)
    task_total = torch.zeros(1, dtype=torch.int64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    if n_tasks:
        task_prefix = torch.empty(n_tasks, dtype=torch.int64, device=dev)
        ws_bytes = lib().skm_ann_sort_workspace(n_tasks)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(lib().skm_ann_sort(_ptr(g_res), nres_g, _ptr(g_off), n_sel, _ptr(tab.lut), tab.nsym, int(k), _ptr(t_lo), _ptr(t_hi), _ptr(t_c0),
                                 _ptr(t_c1), _ptr(t_ann), _ptr(out_base), n_tasks, _ptr(keys), _ptr(vals), capacity, _ptr(task_prefix),
                                 _ptr(totals), _ptr(task_total), _ptr(flag), _ptr(ws), ws_bytes, _stream()))
    if rows is not None and want_totals:
        check(lib().skm_rows_colsum(_ptr(rows), rows.shape[0], S, _ptr(totals), _stream()))
    if H:
        # a heavy annotation's block starts after its own heavy predecessors and after the task entries of smaller annotations
        if n_tasks:
            tp = torch.cat([task_prefix, task_total])
            row_dst = (heavy_cum[:-1] + tp[torch.searchsorted(t_ann, heavy)]).contiguous()
        else:
            row_dst = heavy_cum[:-1].contiguous()
        check(lib().skm_rows_emit(_ptr(rows), H, S, _ptr(blk_off), _ptr(row_dst), _ptr(heavy.contiguous()), _ptr(keys), _ptr(vals), capacity, _stream()))
    n_task_entries, bad = torch.cat([task_total, flag.to(torch.int64)]).tolist()          # one sync
    if bad:
        return learn_sparse_hybrid(batch, alphabet, k, ann_id, n_ann, want_totals=want_totals)
    total = heavy_total + n_task_entries
    keys, vals = keys[:total], vals[:total]
    return (keys, vals, totals) if want_totals else (keys, vals)


def learn_sparse_with_totals(batch: SequenceBatch, alphabet, k: int, ann_id: torch.Tensor, n_ann: int, method: Optional[str] = None):
    """learn_sparse + the Totals row (k-mer occurrences over ALL sequences, learn.smk:380) -> (keys, vals, totals int64 [S]).
    method "hybrid" (default): heavy annotations and the unannotated sequences are counted into dense rows, the Totals are
    column sums (learn_sparse_hybrid).  "sorted": everything through the sort; the annotated share of the totals is the
    column sums of the matrix (one atomic per entry instead of one per window) and the unannotated sequences are
    counted window by window (the round-1 path, kept as the cross-check)."""
    tab = alphabet_tables(alphabet, batch.device)
    dev = batch.device
    S = code_space(tab.nsym, k)
    if S > _native.SKM_DENSE_MAX_SPACE:
        raise SkmError(-3, f"code space {tab.nsym}^{k} exceeds the table limit 2^27")
    ann_id = ann_id.to(device=dev, dtype=torch.int32).contiguous()
    method = method or os.environ.get("SKM_LEARN_METHOD", "hybrid")
    if method == "onchip":
        return learn_sparse_onchip(batch, alphabet, k, ann_id, n_ann, want_totals=True)
    if method == "hybrid":
        return learn_sparse_hybrid(batch, alphabet, k, ann_id, n_ann, want_totals=True)
    keys, vals = learn_sparse(batch, alphabet, k, ann_id, n_ann)
    totals = torch.zeros(S, dtype=torch.int64, device=dev)
    check(lib().skm_coo_colsum(_ptr(keys), _ptr(vals), keys.numel(), S, _ptr(totals), _stream()))
    rest = torch.nonzero((ann_id < 0) | (ann_id >= n_ann)).reshape(-1)
    if rest.numel():
        g_res, g_off = gather_sequences(batch, rest)
        off_host = g_off.cpu().numpy()
        sub = SequenceBatch(g_res, g_off, off_host)
        basis_accumulate(sub, alphabet, k, totals, None, 0)
    return keys, vals, totals


@dataclass
class AnnotationCSC:
    """k-mer-major view of an annotation slice [ann_lo, ann_lo + n_ann) of a learned matrix."""
    colptr: torch.Tensor    # int64 [S+1]
    rows: torch.Tensor      # int32 [nnz] annotation index relative to ann_lo
    mvals: torch.Tensor     # int32 [nnz] = M[a, c]
    mnorm2: torch.Tensor    # float64 [n_ann], exact integer sums
    inv_m32: torch.Tensor   # float32 [n_ann] = 1 / ||m_a|| (screening pass)
    max_m: int              # largest entry
    n_ann: int
    ann_lo: int
    S: int
    packed: Optional[torch.Tensor] = None   # int32 (uint32 pattern) [nnz]: value << 16 | row, when n_ann <= 65536 and max_m < 65536


def csc_build(keys: torch.Tensor, vals: torch.Tensor, S: int, n_ann: int, ann_lo: int = 0, pack: bool = True) -> AnnotationCSC:
    """CSC of the annotation slice [ann_lo, ann_lo + n_ann) of a sorted COO matrix.  pack: also keep the entries as one
    32-bit word each (value << 16 | annotation) when they fit — apply_sparse then reads half the bytes."""
    dev = _require_cuda(keys.device)
    if keys.numel():
        # the slice is contiguous because the list is sorted by ann * S + code
        bounds = torch.tensor([ann_lo * S, (ann_lo + n_ann) * S], dtype=torch.int64, device=dev)
        i0, i1 = torch.searchsorted(keys, bounds).tolist()
        keys = (keys[i0:i1] - ann_lo * S).contiguous()
        vals = vals[i0:i1].contiguous()
    nnz = keys.numel()
    ws_bytes = lib().skm_csc_build_workspace(nnz, n_ann)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    colptr = torch.empty(S + 1, dtype=torch.int64, device=dev)
    rows = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    mvals = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    mn2 = torch.zeros(max(n_ann, 1), dtype=torch.float64, device=dev)
    inv32 = torch.zeros(max(n_ann, 1), dtype=torch.float32, device=dev)
    max_m = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().skm_csc_build(_ptr(keys), _ptr(vals), nnz, int(S), int(n_ann), _ptr(colptr), _ptr(rows), _ptr(mvals), _ptr(mn2),
                              _ptr(inv32), _ptr(max_m), _ptr(ws), ws_bytes, _stream()))
    mx = int(max_m.item())
    if mx >= 2 ** 31:
        raise SkmError(-3, f"csc_build: an annotation count of {mx} does not fit the 32-bit CSC values")
    csc = AnnotationCSC(colptr, rows[:nnz], mvals[:nnz], mn2[:n_ann], inv32[:n_ann], mx, int(n_ann), int(ann_lo), int(S))
    if pack and nnz and n_ann <= 65536 and mx < 65536:
        csc.packed = torch.empty(nnz, dtype=torch.int32, device=dev)
        check(lib().skm_csc_pack(_ptr(csc.rows), _ptr(csc.mvals), nnz, _ptr(csc.packed), _stream()))
    return csc


SPARSE_MAX_ANN = 50 * 1024          # annotations per apply_sparse call with 32-bit accumulators (half with 64-bit)


def apply_sparse(rowptr: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, csc: AnnotationCSC,
                 max_row_total: Optional[int] = None, use_packed: bool = True) -> ApplyResult:
    """SpMM scoring of CSR queries (codes + counts, count_csr with basis=None) against one annotation slice.
    max_row_total bounds the sum of a query's counts (default: the largest row total, one reduction); it decides
    whether the exact integer dots fit 32-bit accumulators."""
    dev = _require_cuda(rowptr.device)
    nq = rowptr.numel() - 1
    if nq == 0 or csc.n_ann == 0:               # no candidates: -1 / NaN, like apply_dense
        return ApplyResult(torch.full((nq,), -1, dtype=torch.int32, device=dev), torch.full((nq,), -1, dtype=torch.int32, device=dev),
                           torch.full((nq,), float("nan"), dtype=torch.float64, device=dev),
                           torch.full((nq,), float("nan"), dtype=torch.float64, device=dev), None)
    top1 = torch.empty(nq, dtype=torch.int32, device=dev)
    top2 = torch.empty(nq, dtype=torch.int32, device=dev)
    s1 = torch.empty(nq, dtype=torch.float64, device=dev)
    s2 = torch.empty(nq, dtype=torch.float64, device=dev)
    if max_row_total is None:
        max_row_total = query_row_total_max(rowptr, vals)
    acc_bits = sparse_acc_bits(csc.max_m, max_row_total)
    if csc.n_ann > sparse_max_ann(acc_bits):
        raise SkmError(-3, f"apply_sparse: {csc.n_ann} annotations with {acc_bits}-bit accumulators exceed one pass "
                           f"({sparse_max_ann(acc_bits)}); use apply_sparse_tiled")
    check(lib().skm_apply_sparse(_ptr(rowptr), _ptr(cols), _ptr(vals), nq, _ptr(csc.colptr), _ptr(csc.rows), _ptr(csc.mvals),
                                 _ptr(csc.packed) if use_packed else None, _ptr(csc.mnorm2), _ptr(csc.inv_m32), csc.n_ann, acc_bits, _ptr(top1), _ptr(top2), _ptr(s1), _ptr(s2),
                                 None, _stream()))
    return ApplyResult(top1, top2, s1, s2, None)


def query_row_total_max(rowptr: torch.Tensor, vals: torch.Tensor) -> int:
    """Largest sum of counts over a CSR row (= number of valid windows of the longest query)."""
    if vals.numel() == 0:
        return 0
    csum = torch.zeros(vals.numel() + 1, dtype=torch.int64, device=vals.device)
    torch.cumsum(vals, 0, out=csum[1:])
    return int((csum[rowptr[1:]] - csum[rowptr[:-1]]).max().item())


def sparse_acc_bits(max_m: int, max_row_total: int) -> int:
    """Width of the exact integer accumulators of apply_sparse: a dot is at most max_m * (sum of the query's counts)."""
    return 32 if int(max_m) * max(int(max_row_total), 1) < 2 ** 32 else 64


def sparse_max_ann(acc_bits: int) -> int:
    """Annotations one apply_sparse pass holds in shared memory (200 KB of accumulators)."""
    return SPARSE_MAX_ANN if acc_bits == 32 else SPARSE_MAX_ANN // 2


def apply_sparse_tiled(rowptr, cols, vals, keys, mvals, S: int, n_ann: int, tile: Optional[int] = None) -> ApplyResult:
    """All annotations of a sorted COO matrix, `tile` at a time, merged with the top-2 merge (the same fan-in as the
    multi-GPU annotation sharding).  The tile defaults to what one pass can hold, which depends on the accumulator
    width: 51,200 annotations with 32-bit accumulators, 25,600 when max(M) * (largest query total) needs 64 bits."""
    dev = rowptr.device
    nq = rowptr.numel() - 1
    if n_ann <= 0 or nq <= 0:
        return ApplyResult(torch.full((nq,), -1, dtype=torch.int32, device=dev), torch.full((nq,), -1, dtype=torch.int32, device=dev),
                           torch.full((nq,), float("nan"), dtype=torch.float64, device=dev),
                           torch.full((nq,), float("nan"), dtype=torch.float64, device=dev), None)
    row_total = query_row_total_max(rowptr, vals)
    max_m = int(mvals.max().item()) if mvals.numel() else 0
    cap = sparse_max_ann(sparse_acc_bits(max_m, row_total))
    tile = cap if tile is None else max(1, min(int(tile), cap))
    idxs, scs = [], []
    for a0 in range(0, n_ann, tile):
        na = min(tile, n_ann - a0)
        r = apply_sparse(rowptr, cols, vals, csc_build(keys, mvals, S, na, a0), row_total)
        if n_ann <= tile:
            return r
        i = torch.stack([r.top1.to(torch.int64), r.top2.to(torch.int64)])
        idxs.append(torch.where(i >= 0, i + a0, i))
        scs.append(torch.stack([r.score1, r.score2]))
    return merge_top2(torch.stack(idxs), torch.stack(scs))


# ---------------------------------------------------------------------------
# wide path: code spaces beyond the table limit (nsym^k > 2^27, up to 2^64 - 1) — sort-based
# ---------------------------------------------------------------------------
WIDE_MAX_CHUNK_RES = (1 << 30) - 16


@dataclass
class WideBasis:
    """Basis over a code space too large for tables: codes in first-occurrence order plus the
    ascending code list with the column of each entry (binary-search lookup)."""
    alphabet: str
    k: int
    symbols: str
    codes: torch.Tensor             # int64 (uint64 pattern) [K], first-occurrence order
    counts: Optional[torch.Tensor]  # int64 [K]
    sorted_codes: torch.Tensor      # int64 (uint64 pattern) [K], ascending (unsigned)
    col_of_sorted: torch.Tensor     # int32 [K]
    K: int

    def codes_host(self) -> np.ndarray:
        return self.codes.cpu().numpy().view(np.uint64)

    def kmers(self) -> np.ndarray:
        return decode_kmers(self.codes_host(), self.symbols, self.k)


def basis_table_local(batch: SequenceBatch, alphabet: AlphabetT, k: int, res_base: int = 0,
                      max_chunk_res: int = WIDE_MAX_CHUNK_RES) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, bool]:
    """(codes, counts, first) of a shard: distinct window codes sorted by code, their occurrence counts and
    the global position of their first occurrence.  Returns a 4th item: True when the table is one merged
    list (a single chunk), False when it is the concatenation of several chunks' tables."""
    tab = alphabet_tables(alphabet, batch.device)
    dev = batch.device
    parts = []
    for lo, hi in _chunks_by_residues(batch.offsets_host, max_chunk_res):
        sub = batch if (lo == 0 and hi == batch.n) else _sub_batch(batch, lo, hi)
        shift = 0 if sub is batch else (int(batch.offsets_host[lo]) & ~15)
        n_cap = max(sub.nres, 1)
        ws_bytes = lib().skm_basis_sorted_local_workspace(sub.nres)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        codes = torch.empty(n_cap, dtype=torch.int64, device=dev)
        counts = torch.empty(n_cap, dtype=torch.int64, device=dev)
        first = torch.empty(n_cap, dtype=torch.int64, device=dev)
        dn = torch.zeros(1, dtype=torch.int64, device=dev)
        check(lib().skm_basis_sorted_local(_ptr(sub.residues), sub.nres, _ptr(sub.offsets), sub.n, _ptr(tab.lut), tab.nsym, int(k),
                                           int(res_base) + shift, _ptr(codes), _ptr(counts), _ptr(first), _ptr(dn), _ptr(ws),
                                           ws_bytes, _stream()))
        m = int(dn.item())
        parts.append((codes[:m].clone(), counts[:m].clone(), first[:m].clone()))
        del ws, codes, counts, first
    if not parts:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, z.clone(), z.clone(), True
    if len(parts) == 1:
        return parts[0] + (True,)
    return (torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), torch.cat([p[2] for p in parts]), False)


def basis_table_finalize(alphabet: AlphabetT, k: int, codes: torch.Tensor, counts: torch.Tensor, first: torch.Tensor,
                         merged: bool, min_filter: int = 0, first_bound: int = 0) -> WideBasis:
    """Tables of one or more shards -> the basis in the reference's order (kmerize.smk:89-104).
    first_bound: exclusive upper bound of the first positions (total residues of all shards), 0 = unknown."""
    dev = _require_cuda(codes.device)
    tab = alphabet_tables(alphabet, dev)
    n = codes.numel()
    codes, counts, first = codes.contiguous(), counts.contiguous(), first.contiguous()
    ws_bytes = lib().skm_basis_sorted_finalize_workspace(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cap = max(n, 1)
    basis = torch.empty(cap, dtype=torch.int64, device=dev)
    bcnt = torch.empty(cap, dtype=torch.int64, device=dev)
    scodes = torch.empty(cap, dtype=torch.int64, device=dev)
    scol = torch.empty(cap, dtype=torch.int32, device=dev)
    dK = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib().skm_basis_sorted_finalize(_ptr(codes), _ptr(counts), _ptr(first), n, 1 if merged else 0, int(min_filter), int(first_bound),
                                          _ptr(basis), _ptr(bcnt), _ptr(scodes), _ptr(scol), _ptr(dK), _ptr(ws), ws_bytes, _stream()))
    K = int(dK.item())
    return WideBasis(tab.name, int(k), tab.symbols, basis[:K].clone(), bcnt[:K].clone(), scodes[:K].clone(), scol[:K].clone(), K)


def build_basis_wide(batch: SequenceBatch, alphabet: AlphabetT, k: int, min_filter: int = 0) -> WideBasis:
    """Pass 1 of the vectorize rule (kmerize.smk:89-104) for any nsym^k <= 2^64 - 1: sort-based."""
    codes, counts, first, merged = basis_table_local(batch, alphabet, k, 0)
    return basis_table_finalize(alphabet, k, codes, counts, first, merged, min_filter, first_bound=batch.nres + 16)


def wide_basis_from_codes(codes: np.ndarray, alphabet: AlphabetT, k: int, device=None) -> WideBasis:
    """Basis whose column j holds code codes[j] (uint64, distinct): a supplied / learned k-mer list."""
    dev = _require_cuda(device)
    tab = alphabet_tables(alphabet, dev)
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    order = np.argsort(codes, kind="stable")
    d_codes = torch.from_numpy(codes.view(np.int64)).to(dev)
    return WideBasis(tab.name, int(k), tab.symbols, d_codes, None, torch.from_numpy(codes[order].view(np.int64)).to(dev),
                     torch.from_numpy(order.astype(np.int32)).to(dev), len(codes))


def codes_to_columns(codes: torch.Tensor, basis: WideBasis) -> torch.Tensor:
    """Basis column of every code (int32, -1 = not in the basis)."""
    dev = _require_cuda(codes.device)
    codes = codes.contiguous()
    out = torch.empty(codes.numel(), dtype=torch.int32, device=dev)
    check(lib().skm_codes_to_columns(_ptr(codes), codes.numel(), _ptr(basis.sorted_codes), _ptr(basis.col_of_sorted), basis.K,
                                     _ptr(out), _stream()))
    return out


def count_csr_wide(batch: SequenceBatch, alphabet: AlphabetT, k: int, basis: Optional[WideBasis] = None,
                   max_chunk_res: int = WIDE_MAX_CHUNK_RES, method: str = "warp"):
    """Per-sequence distinct k-mers and their counts for any nsym^k <= 2^64 - 1.
    Returns (rowptr int64 [N+1], codes int64 (uint64 pattern) [nnz], cols int32 [nnz] or None, vals int32 [nnz]);
    a row's entries are ordered by code; with a basis, entries outside it are dropped and cols holds columns."""
    tab = alphabet_tables(alphabet, batch.device)
    dev = batch.device
    if basis is not None and basis.K == 0:           # nothing can match an empty basis
        return (torch.zeros(batch.n + 1, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev),
                torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev))
    rowptrs, codes_p, cols_p, vals_p = [], [], [], []
    base_nnz = 0
    for lo, hi in _chunks_by_residues(batch.offsets_host, max_chunk_res):
        sub = batch if (lo == 0 and hi == batch.n) else _sub_batch(batch, lo, hi)
        cap = max(sub.nres, 1)
        rowptr = torch.empty(sub.n + 1, dtype=torch.int64, device=dev)
        codes = torch.empty(cap, dtype=torch.int64, device=dev)
        cols = torch.empty(cap, dtype=torch.int32, device=dev) if basis is not None else None
        vals = torch.empty(cap, dtype=torch.int32, device=dev)
        b_codes = None if basis is None else _ptr(basis.sorted_codes)
        b_cols = None if basis is None else _ptr(basis.col_of_sorted)
        b_K = 0 if basis is None else basis.K
        if method == "segsort":          # the first implementation (window keys in HBM + cub::DeviceSegmentedSort), a cross-check
            ws_bytes = lib().skm_count_csr_wide_workspace(sub.nres, sub.n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib().skm_count_csr_wide(_ptr(sub.residues), sub.nres, _ptr(sub.offsets), sub.n, _ptr(tab.lut), tab.nsym, int(k),
                                           b_codes, b_cols, b_K, _ptr(rowptr), _ptr(codes), _ptr(cols), _ptr(vals), _ptr(ws), ws_bytes, _stream()))
        else:                            # a warp sorts one sequence's keys in shared memory
            max_len = int(np.diff(sub.offsets_host).max()) if sub.n else 0
            ws_bytes = lib().skm_count_csr_sorted_workspace(sub.nres, sub.n, max_len, 64)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib().skm_count_csr_sorted(_ptr(sub.residues), sub.nres, _ptr(sub.offsets), sub.n, _ptr(tab.lut), tab.nsym, int(k), 64,
                                             None, 0, b_codes, b_cols, b_K, max_len, _ptr(rowptr), _ptr(codes), _ptr(cols), _ptr(vals),
                                             _ptr(ws), ws_bytes, _stream()))
        nnz = int(rowptr[-1].item()) if sub.n else 0
        rowptrs.append(rowptr[(1 if rowptrs else 0):] + base_nnz)
        codes_p.append(codes[:nnz].clone())
        vals_p.append(vals[:nnz].clone())
        if cols is not None:
            cols_p.append(cols[:nnz].clone())
        base_nnz += nnz
        del ws, codes, cols, vals
    if not rowptrs:
        z = torch.zeros(1, dtype=torch.int64, device=dev)
        e64 = torch.zeros(0, dtype=torch.int64, device=dev)
        e32 = torch.zeros(0, dtype=torch.int32, device=dev)
        return z, e64, (e32 if basis is not None else None), e32.clone()
    cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs))
    return cat(rowptrs), cat(codes_p), (cat(cols_p) if basis is not None else None), cat(vals_p)


def build_basis_wide_distributed(batch: SequenceBatch, alphabet, k: int, min_filter: int = 0,
                                 res_base: Optional[int] = None) -> WideBasis:
    """The wide basis of the CONCATENATION of all ranks' shards, identical on every rank: local tables with
    global first positions, all_gather of the tables, one finalisation (sort by code, sum / min, order)."""
    from . import dist as D

    total = 0
    if res_base is None:
        res_base, total = D.exclusive_prefix(batch.nres, batch.device)
    codes, counts, first, merged = basis_table_local(batch, alphabet, k, res_base)
    _, w = D.world()
    if w > 1:
        codes, counts, first = D.allgather_tables(codes, counts, first)
        merged = False
    return basis_table_finalize(alphabet, k, codes, counts, first, merged, min_filter, first_bound=(total + 16 * w) if total else 0)


DENSE_ROW_MAX_K = 4096          # widest basis whose per-sequence counts are kept as dense rows by vectorize()


@dataclass
class Vectorized:
    """Result of vectorize(): the basis and the per-sequence counts in the layout the basis size calls for."""
    path: str                               # "dense" | "csr" | "wide"
    basis: "Basis | WideBasis"
    counts: Optional[torch.Tensor] = None   # dense: int32 [N, K]
    rowptr: Optional[torch.Tensor] = None   # csr / wide: int64 [N+1]
    cols: Optional[torch.Tensor] = None     #             int32 [nnz] basis column of every entry
    vals: Optional[torch.Tensor] = None     #             int32 [nnz]
    codes: Optional[torch.Tensor] = None    # wide only: int64 (uint64 pattern) [nnz]

    @property
    def K(self) -> int:
        return self.basis.K

    @property
    def nnz(self) -> int:
        return int(self.vals.numel()) if self.vals is not None else int((self.counts != 0).sum().item())


def vectorize(batch: SequenceBatch, alphabet: AlphabetT, k: int, min_filter: int = 0,
              dense_max_K: int = DENSE_ROW_MAX_K) -> Vectorized:
    """Both passes of the vectorize rule (kmerize.smk:85-120) for ANY alphabet / k with nsym^k <= 2^64 - 1:
    first-occurrence basis + per-sequence counts.  Code spaces up to 2^27 use the table kernels (dense rows while
    the basis has at most `dense_max_K` columns, CSR beyond); larger ones the sort-based wide path."""
    tab = alphabet_tables(alphabet, batch.device)
    S = code_space(tab.nsym, k)
    if S <= _native.SKM_DENSE_MAX_SPACE:
        basis = build_basis(batch, alphabet, k, min_filter)
        if basis.K <= dense_max_K:
            return Vectorized("dense", basis, counts=count_dense(batch, alphabet, k, basis))
        parts, base = [], 0
        for lo, hi in _chunks_by_residues(batch.offsets_host, (1 << 30) - 16):
            sub = batch if (lo == 0 and hi == batch.n) else _sub_batch(batch, lo, hi)
            rp, c, v = count_csr(sub, alphabet, k, basis)
            parts.append((rp[(1 if parts else 0):] + base, c, v))
            base += int(v.numel())
        if not parts:
            z32 = torch.zeros(0, dtype=torch.int32, device=batch.device)
            return Vectorized("csr", basis, rowptr=torch.zeros(1, dtype=torch.int64, device=batch.device), cols=z32, vals=z32.clone())
        cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs))
        return Vectorized("csr", basis, rowptr=cat([p[0] for p in parts]), cols=cat([p[1] for p in parts]), vals=cat([p[2] for p in parts]))
    basis = build_basis_wide(batch, alphabet, k, min_filter)
    rowptr, codes, cols, vals = count_csr_wide(batch, alphabet, k, basis)
    return Vectorized("wide", basis, rowptr=rowptr, cols=cols, vals=vals, codes=codes)


# ---------------------------------------------------------------------------
# multi-GPU compositions (one process per GPU; collectives in dist.py)
# ---------------------------------------------------------------------------
def build_basis_distributed(batch: SequenceBatch, alphabet, k: int, min_filter: int = 0, res_base: Optional[int] = None) -> Basis:
    """The basis of the CONCATENATION of all ranks' shards (rank order), identical on every rank:
    local tables with global first positions, all_reduce(SUM / MIN), local finalisation."""
    from . import dist as D

    tab = alphabet_tables(alphabet, batch.device)
    S = code_space(tab.nsym, k)
    if S > _native.SKM_DENSE_MAX_SPACE:
        raise SkmError(-3, f"code space {tab.nsym}^{k} exceeds the table limit 2^27")
    if res_base is None:
        res_base, _ = D.exclusive_prefix(batch.nres, batch.device)
    count, first = basis_tables(S, batch.device)
    basis_accumulate(batch, alphabet, k, count, first, res_base)
    D.allreduce_basis_tables(count, first)
    return basis_finalize(alphabet, k, count, first, min_filter)


def learn_dense_distributed(batch: SequenceBatch, alphabet, k: int, basis: Optional[Basis], ann_id: torch.Tensor,
                            n_ann: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """learn_dense on the local shard followed by the NCCL sum over ranks (the reference's
    serial Merge step, learn.smk:467-494); every rank ends with the full matrix."""
    from . import dist as D

    M, totals = learn_dense(batch, alphabet, k, basis, ann_id, n_ann)
    D.allreduce_sum_(M, totals)
    return M, totals


def apply_dense_annotation_sharded(Q: torch.Tensor, M_local: torch.Tensor, ann_base: int,
                                   qnorm2: Optional[torch.Tensor] = None) -> ApplyResult:
    """Every rank scores all queries against ITS rows of the annotation matrix; the per-shard
    top-2 are all-gathered and merged (identical result on every rank)."""
    from . import dist as D

    r = apply_dense(Q, M_local, qnorm2=qnorm2)
    idx, sc = D.allgather_top2(r.top1, r.top2, r.score1, r.score2, ann_base)
    return merge_top2(idx, sc)


def peer_exchange_enabled(t: torch.Tensor) -> bool:
    """The sparse-learn fan-in goes through NVLink peer memory when the ranks are CUDA processes of one node (NCCL
    backend); SKM_EXCHANGE=nccl selects the all_to_all_single baseline."""
    import torch.distributed as dist
    from . import dist as D
    return (t.is_cuda and dist.is_initialized() and dist.get_backend() == "nccl" and dist.get_world_size() <= 16
            and os.environ.get("SKM_EXCHANGE", "peer") != "nccl" and not D.peer_buffers().disabled)


def exchange_coo_by_annotation(keys: torch.Tensor, vals: torch.Tensor, S: int, n_ann: int,
                               balance: bool = True) -> Tuple[torch.Tensor, torch.Tensor, Tuple[int, int]]:
    """Multi-GPU fan-in of sparse learn: annotations are split into W contiguous ranges, rank r ends up with the merged
    rows [a_lo, a_hi) of the matrix — the annotation sharding apply uses.  Returns (keys, vals, (a_lo, a_hi)).

    balance=True cuts the ranges by ENTRY count (one all_reduce of the per-annotation entry histogram): family sizes
    are Zipf-distributed, so equal id ranges would send ~90 % of all entries to rank 0.  Cuts stay at annotation
    boundaries (a rank owns whole rows), so a single annotation larger than 1/W of the matrix still limits the balance."""
    from . import dist as D

    rank, w = D.world()
    if w == 1:
        return keys, vals, (0, n_ann)
    if balance:
        ann_bounds = D.balanced_annotation_bounds(keys, S, n_ann)
    else:
        ann_bounds = [n_ann * r // w for r in range(w)] + [n_ann]
    key_bounds = [a * int(S) for a in ann_bounds]
    count_bits = 64 - max(int(n_ann) * int(S) - 1, 1).bit_length()
    if peer_exchange_enabled(keys) and count_bits >= 16:
        # fused pack + all_to_all over NVLink peer memory (skm_coo_pack_push); the histogram all_reduce above (or the
        # barrier below) is the fence that says no rank still reads its receive buffer
        if not balance:
            D.barrier()
        pushed = D.push_coo_by_key_range(keys, vals, key_bounds, count_bits)
        if pushed is not None:
            ptr, runs, flag = pushed
            k3, v3, dn = coo_merge_runs_packed(ptr, runs, count_bits, keys.device)
            m, bad = torch.cat([dn, flag.to(torch.int64)]).tolist()
            if not bad:
                return k3[:m].clone(), v3[:m].clone(), (ann_bounds[rank], ann_bounds[rank + 1])
        # a count beyond count_bits somewhere: the unpacked NCCL exchange below
    k2, v2, runs = D.alltoall_coo_by_key_range(keys, vals, key_bounds, return_runs=True)
    k3, v3 = coo_merge_runs(k2, v2, runs)           # W sorted runs, one per sender: merge tree + reduce-by-key
    return k3, v3, (ann_bounds[rank], ann_bounds[rank + 1])
