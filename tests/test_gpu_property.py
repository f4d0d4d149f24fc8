"""Property-based GPU parity (hypothesis): adversarial sequence sets — empties, runs of invalid residues, trailing
stars, lengths around the 16-byte staging granule and the k-1 halo, sequences starting at every alignment — through every
counting path (dense rows, table CSR by warp sort and by segmented sort, wide CSR, sparse learn) against the oracle."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import skm_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from snekmer_b200 import engine as E

ALPHA = "ACDEFGHIKLMNPQRSTVWY"
piece = st.one_of(
    st.text(alphabet=ALPHA, min_size=0, max_size=40),
    st.text(alphabet=ALPHA + "XBZ*ac", min_size=0, max_size=24),
    st.sampled_from(["", "*", "**", "X", "A" * 17, "AC" * 16, "ACDEFGHIKLMNPQR", "ACDEFGHIKLMNPQRS", "ACDEFGHIKLMNPQRST", "K" * 64]),
    st.builds(lambda c, n: c * n, st.sampled_from(list(ALPHA)), st.integers(0, 130)),
)
seqs_strategy = st.lists(st.builds(lambda parts: "".join(parts), st.lists(piece, min_size=0, max_size=6)), min_size=1, max_size=60)
case = st.tuples(st.sampled_from([0, 2, 3, 5, None, "ptm"]), st.integers(1, 9))


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(seqs=seqs_strategy, ak=case, mf=st.integers(0, 2))
def test_counting_paths_agree_with_oracle(seqs, ak, mf):
    a, k = ak
    lut, syms = O.build_lut(a)
    nsym = len(syms)
    res, offs = O.pack(seqs)
    si, pos, code, valid = O.window_codes(res, offs, lut, nsym, k)
    want_b, want_cnt = O.basis_codes(si, pos, code, valid, mf)
    want_csr = O.count_csr(si, code, valid, len(seqs))
    batch = E.SequenceBatch.from_strings(seqs)
    S = nsym ** k
    # wide path (any code space)
    wb = E.build_basis_wide(batch, a, k, mf)
    assert np.array_equal(wb.codes_host(), want_b) and np.array_equal(wb.counts.cpu().numpy(), want_cnt)
    for method in ("warp", "segsort"):
        rp, codes, _, vals = E.count_csr_wide(batch, a, k, None, method=method)
        assert np.array_equal(rp.cpu().numpy(), want_csr[0]) and np.array_equal(codes.cpu().numpy().view(np.uint64), want_csr[1])
        assert np.array_equal(vals.cpu().numpy(), want_csr[2])
    if S <= 2 ** 27:
        tb = E.build_basis(batch, a, k, mf)
        assert np.array_equal(tb.codes_host(), want_b) and np.array_equal(tb.counts.cpu().numpy(), want_cnt)
        if tb.K * len(seqs) < 5e6 and tb.K * 4 < 60000:
            dense = O.count_matrix(si, code, valid, len(seqs), want_b)
            assert np.array_equal(E.count_dense(batch, a, k, tb).cpu().numpy(), dense)
        r1 = E.count_csr(batch, a, k, tb, method="warp")
        r2 = E.count_csr(batch, a, k, tb, method="segsort")
        assert all(torch.equal(x, y) for x, y in zip(r1, r2))
        assert int(r1[2].sum().item()) == int((np.isin(code, want_b) & valid).sum())
        # sparse learn: grouped == global, and row sums == per-annotation window counts
        n_ann = 5
        ann = (np.arange(len(seqs)) % (n_ann + 1) - 1).astype(np.int32)
        k1, v1 = E.learn_sparse(batch, a, k, torch.from_numpy(ann), n_ann)
        k2, v2 = E.learn_sparse(batch, a, k, torch.from_numpy(ann), n_ann, method="global")
        assert torch.equal(k1, k2) and torch.equal(v1, v2)
        per_ann = np.zeros(n_ann, np.int64)
        np.add.at(per_ann, ann[si[valid]][ann[si[valid]] >= 0], 1)
        got = np.zeros(n_ann, np.int64)
        np.add.at(got, (k1.cpu().numpy() // S), v1.cpu().numpy())
        assert np.array_equal(got, per_ann)
