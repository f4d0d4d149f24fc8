"""GPU parity tests of the wide (sort-based) vectorisation path: code spaces beyond the 2^27 table limit up to
2^64 - 1 (the alphabet / k sweep, BASELINE config C5), against the CPU oracle.  Bit-exact: basis codes and their
order, occurrence counts, per-sequence (code, count) lists, basis columns."""
import numpy as np
import pytest
import torch

from oracle import skm_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from snekmer_b200 import engine as E


def _rand_seqs(rng, n, lo=0, hi=600, p_x=0.002):
    aa = np.array(list("ACDEFGHIKLMNPQRSTVWYXBZUO*acd"))
    p = np.array([1.0] * 20 + [p_x * 20] * 9)
    p /= p.sum()
    return ["".join(rng.choice(aa, size=int(rng.integers(lo, hi)), p=p)) for _ in range(n)]


def _oracle(seqs, a, k):
    lut, syms = O.build_lut(a)
    res, offs = O.pack(seqs)
    return O.window_codes(res, offs, lut, len(syms), k), offs


def _u64(t):
    return t.cpu().numpy().view(np.uint64)


def _check_csr(got, want, n):
    rowptr, codes, vals = got
    wr, wc, wv = want
    assert np.array_equal(rowptr.cpu().numpy(), wr)
    assert np.array_equal(_u64(codes), wc)
    assert np.array_equal(vals.cpu().numpy(), wv)


# (alphabet, k): small spaces (the wide path must agree with the table path's oracle too), 2^27..2^32, > 2^32,
# power-of-two spaces (hydro 2^k: the all-ones low bits of the invalid key tie with the largest code), near 2^64
CASES = [(5, 3), (0, 5), (2, 8), (5, 9), (None, 7), (None, 8), (None, 14), (0, 32), (0, 63), ("ptm", 12), (1, 22), (3, 40)]


@pytest.mark.parametrize("a,k", CASES)
@pytest.mark.parametrize("mf", [0, 2])
def test_wide_basis_matches_oracle(a, k, mf):
    rng = np.random.default_rng(k * 17 + mf)
    seqs = _rand_seqs(rng, 400, 0, 500)
    (si, pos, code, valid), _ = _oracle(seqs, a, k)
    want, want_cnt = O.basis_codes(si, pos, code, valid, mf)
    batch = E.SequenceBatch.from_strings(seqs)
    b = E.build_basis_wide(batch, a, k, mf)
    assert b.K == len(want)
    assert np.array_equal(b.codes_host(), want)
    assert np.array_equal(b.counts.cpu().numpy(), want_cnt)
    # lookup structure: ascending codes, column of each
    sc = _u64(b.sorted_codes)
    assert np.array_equal(sc, np.sort(want))
    assert np.array_equal(want[b.col_of_sorted.cpu().numpy()], sc)
    # chunked shard == one chunk (tables of chunks merged by the finalisation)
    b2 = E.basis_table_finalize(a, k, *E.basis_table_local(batch, a, k, 0, max_chunk_res=20000)[:3], merged=False, min_filter=mf)
    assert np.array_equal(b2.codes_host(), want) and np.array_equal(b2.counts.cpu().numpy(), want_cnt)


@pytest.mark.parametrize("a,k", CASES)
def test_wide_csr_matches_oracle(a, k):
    rng = np.random.default_rng(k * 29 + 1)
    seqs = _rand_seqs(rng, 300, 0, 500) + ["", "A", "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA" * 3]
    (si, pos, code, valid), _ = _oracle(seqs, a, k)
    batch = E.SequenceBatch.from_strings(seqs)
    rowptr, codes, cols, vals = E.count_csr_wide(batch, a, k)
    assert cols is None
    _check_csr((rowptr, codes, vals), O.count_csr(si, code, valid, len(seqs)), len(seqs))
    # chunked
    r2, c2, _, v2 = E.count_csr_wide(batch, a, k, max_chunk_res=15000)
    _check_csr((r2, c2, v2), O.count_csr(si, code, valid, len(seqs)), len(seqs))
    # over a basis with a min_filter: entries outside it are dropped, columns = basis order
    want_b, _ = O.basis_codes(si, pos, code, valid, 1)
    basis = E.build_basis_wide(batch, a, k, 1)
    rowptr, codes, cols, vals = E.count_csr_wide(batch, a, k, basis)
    inb = np.isin(code, want_b) & valid
    _check_csr((rowptr, codes, vals), O.count_csr(si, code, inb, len(seqs)), len(seqs))
    assert np.array_equal(want_b[cols.cpu().numpy()], _u64(codes))
    # the lookup on its own, with codes the basis does not hold
    probe = np.concatenate([want_b[:50], want_b[:50] + np.uint64(1), np.array([0, 2 ** 64 - 1], dtype=np.uint64)])
    got = E.codes_to_columns(torch.from_numpy(probe.view(np.int64)).cuda(), basis).cpu().numpy()
    pos_in = {int(c): i for i, c in enumerate(want_b)}
    assert got.tolist() == [pos_in.get(int(c), -1) for c in probe]


def test_wide_supplied_basis_and_dense_agreement():
    """A supplied k-mer list (basis.txt branch) through the wide lookup == the table path's dense counts."""
    rng = np.random.default_rng(3)
    seqs = _rand_seqs(rng, 500, 0, 400)
    a, k = 5, 3
    batch = E.SequenceBatch.from_strings(seqs)
    tb = E.build_basis(batch, a, k, 0)
    dense = E.count_dense(batch, a, k, tb).cpu().numpy()
    wb = E.wide_basis_from_codes(tb.codes_host(), a, k)
    rowptr, codes, cols, vals = E.count_csr_wide(batch, a, k, wb)
    rp = rowptr.cpu().numpy()
    rebuilt = np.zeros_like(dense)
    rows = np.repeat(np.arange(len(seqs)), np.diff(rp))
    rebuilt[rows, cols.cpu().numpy()] = vals.cpu().numpy()
    assert np.array_equal(rebuilt, dense)
    assert np.array_equal(E.build_basis_wide(batch, a, k, 0).codes_host(), tb.codes_host())


def test_wide_empty_and_errors():
    batch = E.SequenceBatch.from_strings([])
    b = E.build_basis_wide(batch, None, 14, 0)
    assert b.K == 0
    rowptr, codes, cols, vals = E.count_csr_wide(batch, None, 14)
    assert rowptr.tolist() == [0] and codes.numel() == 0
    batch = E.SequenceBatch.from_strings(["ACD", "XXXX", ""])
    b = E.build_basis_wide(batch, None, 14, 0)
    assert b.K == 0
    rowptr, codes, cols, vals = E.count_csr_wide(batch, None, 14, b)
    assert rowptr.tolist() == [0, 0, 0, 0]
    with pytest.raises(E.SkmError):       # 2^64 codes collide with the invalid sentinel
        E.build_basis_wide(E.SequenceBatch.from_strings(["AAAA"]), 0, 64, 0)


def test_wide_sweep_properties_full_size():
    """C5 shape at bench size (size-independent properties): counts sum to the number of valid windows, every
    row's codes strictly ascend, basis counts == column sums of the CSR, first-occurrence order is consistent
    with a second run on the reversed half (idempotence of the table merge)."""
    rng = np.random.default_rng(5)
    n = 20000
    lens = np.clip(np.round(rng.lognormal(5.68, 0.6, n)), 30, 5000).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)
    p = np.array([.122, .009, .060, .057, .034, .084, .021, .047, .025, .105, .024, .022, .053, .034, .074, .047, .050, .071, .014, .022, .001])
    res = rng.choice(aa, size=int(offs[-1]), p=p / p.sum())
    batch = E.SequenceBatch.from_packed(res, offs)
    a, k = None, 14
    b = E.build_basis_wide(batch, a, k, 0)
    rowptr, codes, cols, vals = E.count_csr_wide(batch, a, k, b)
    # valid windows = windows without an X
    isx = (res == ord("X")).astype(np.int64)
    cx = np.concatenate([[0], np.cumsum(isx)])
    nvalid = 0
    for i in range(n):
        L = int(lens[i])
        if L >= k:
            s = np.arange(offs[i], offs[i] + L - k + 1)
            nvalid += int(((cx[s + k] - cx[s]) == 0).sum())
    assert int(vals.sum().item()) == nvalid == int(b.counts.sum().item())
    colsum = torch.zeros(b.K, dtype=torch.int64, device="cuda").index_add_(0, cols.long(), vals.long())
    assert torch.equal(colsum, b.counts)
    c = codes.cpu().numpy().view(np.uint64)
    rp = rowptr.cpu().numpy()
    inner = np.ones(len(c), dtype=bool)
    inner[rp[:-1][rp[:-1] < len(c)]] = False
    assert np.all(c[1:][inner[1:]] > c[:-1][inner[1:]])
    # two half tables merged == the whole
    half = n // 2
    t1 = E.basis_table_local(E._sub_batch(batch, 0, half), a, k, 0)
    sub2 = E._sub_batch(batch, half, n)
    t2 = E.basis_table_local(sub2, a, k, int(offs[half]) & ~15)
    bb = E.basis_table_finalize(a, k, torch.cat([t1[0], t2[0]]), torch.cat([t1[1], t2[1]]), torch.cat([t1[2], t2[2]]), False, 0)
    assert torch.equal(bb.codes, b.codes) and torch.equal(bb.counts, b.counts)


@pytest.mark.parametrize("a,k,path", [(5, 3, "dense"), (0, 14, "csr"), (2, 12, "csr"), (None, 7, "wide"), (None, 14, "wide")])
def test_vectorize_dispatch_matches_oracle(a, k, path):
    """engine.vectorize picks dense rows / table CSR / sort-based wide by code-space and basis size; all three
    agree with the oracle's basis (order included) and per-sequence counts."""
    rng = np.random.default_rng(k + 100)
    seqs = _rand_seqs(rng, 400, 0, 500)
    (si, pos, code, valid), _ = _oracle(seqs, a, k)
    want_b, want_cnt = O.basis_codes(si, pos, code, valid, 0)
    batch = E.SequenceBatch.from_strings(seqs)
    v = E.vectorize(batch, a, k, dense_max_K=2048)
    assert v.path == path and v.K == len(want_b)
    assert np.array_equal(v.basis.codes_host(), want_b)
    assert np.array_equal(v.basis.counts.cpu().numpy(), want_cnt)
    dense = O.count_matrix(si, code, valid, len(seqs), want_b) if len(want_b) * len(seqs) < 5e7 else None
    if v.path == "dense":
        assert np.array_equal(v.counts.cpu().numpy(), dense)
    else:
        rp = v.rowptr.cpu().numpy()
        rows = np.repeat(np.arange(len(seqs)), np.diff(rp))
        cols, vals = v.cols.cpu().numpy().astype(np.int64), v.vals.cpu().numpy()
        assert int(vals.sum()) == int(valid.sum())                 # min_filter 0: every valid window is counted
        if dense is not None:
            rebuilt = np.zeros_like(dense)
            rebuilt[rows, cols] = vals
            assert np.array_equal(rebuilt, dense)
        else:
            wr, wc, wv = O.count_csr(si, code, valid, len(seqs))
            assert np.array_equal(rp, wr) and np.array_equal(np.sort(want_b[cols]), np.sort(wc)) and int(vals.sum()) == int(wv.sum())


def _long_mix(rng):
    """Sequences around every buffer boundary of the warp-sort CSR kernel: warp buffer 1024 windows, CTA buffer 8192,
    global scratch beyond; plus low-complexity runs."""
    aa = np.array(list("ACDEFGHIKLMNPQRSTVWYX"))
    lens = [0, 1, 7, 8, 9, 1023, 1024, 1025, 1030, 1031, 1032, 1040, 3000, 8191, 8192, 8199, 8200, 8210, 20000, 33, 350, 350]
    seqs = ["".join(rng.choice(aa, size=n)) for n in lens]
    seqs += ["A" * 2500, "AC" * 700, "ACDEFGHIKL" * 1200, "X" * 50, "ACDX" * 300]
    seqs += _rand_seqs(rng, 200, 0, 600)
    return seqs


@pytest.mark.parametrize("a,k", [(5, 3), (0, 8), (2, 8), (None, 5), (None, 8), (None, 14), (0, 40)])
def test_warp_sort_csr_long_sequences_and_cross_check(a, k):
    rng = np.random.default_rng(k + 7)
    seqs = _long_mix(rng)
    (si, pos, code, valid), _ = _oracle(seqs, a, k)
    want = O.count_csr(si, code, valid, len(seqs))
    batch = E.SequenceBatch.from_strings(seqs)
    nsym = len(O.build_lut(a)[1])
    if nsym ** k < 2 ** 32:
        for method in ("warp", "segsort"):
            rp, cols, vals = E.count_csr(batch, a, k, None, method=method)
            assert np.array_equal(rp.cpu().numpy(), want[0]), method
            assert np.array_equal(cols.cpu().numpy().view(np.uint32).astype(np.uint64), want[1]), method
            assert np.array_equal(vals.cpu().numpy(), want[2]), method
        if nsym ** k <= 2 ** 27:
            tb = E.build_basis(batch, a, k, 1)
            r1 = E.count_csr(batch, a, k, tb, method="warp")
            r2 = E.count_csr(batch, a, k, tb, method="segsort")
            assert all(torch.equal(x, y) for x, y in zip(r1, r2))
    for method in ("warp", "segsort"):
        rp, codes, _, vals = E.count_csr_wide(batch, a, k, None, method=method)
        _check_csr((rp, codes, vals), want, len(seqs))
    wb = E.build_basis_wide(batch, a, k, 1)
    r1 = E.count_csr_wide(batch, a, k, wb, method="warp")
    r2 = E.count_csr_wide(batch, a, k, wb, method="segsort")
    assert all(torch.equal(x, y) for x, y in zip(r1, r2))
    want_b, _ = O.basis_codes(si, pos, code, valid, 1)
    _check_csr((r1[0], r1[1], r1[3]), O.count_csr(si, code, np.isin(code, want_b) & valid, len(seqs)), len(seqs))
    assert np.array_equal(want_b[r1[2].cpu().numpy()], _u64(r1[1]))
