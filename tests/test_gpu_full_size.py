"""BASELINE config C2 at FULL size (1 M synthetic proteins, mean length 350, MIQS k = 3) through size-independent
properties — the oracle is too slow here, the structure of the answer is not: row sums = valid windows per sequence,
column sums = the basis' occurrence counts, the basis is a permutation of the saturated code space in first-occurrence
order, the 16-bit and 32-bit outputs agree, the CSR path carries the same entries, and two runs are bit-identical."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from snekmer_b200 import engine as E


@pytest.mark.timeout(600)
def test_c2_full_size_properties():
    import bench

    n, a, k = 1_000_000, "miqs", 3
    res, off = bench.synth_proteins(n, 2)
    batch = E.SequenceBatch.from_packed(res, off)
    basis = E.build_basis(batch, a, k, 0)
    assert basis.K == 1000                                              # 10^3 codes, all present in 350 M windows
    codes = basis.codes_host().astype(np.int64)
    assert np.array_equal(np.sort(codes), np.arange(1000))
    col = basis.col_of_code.cpu().numpy()
    assert np.array_equal(col[codes], np.arange(1000))
    C = E.count_dense(batch, a, k, basis)
    # valid windows per sequence from the residue buffer: windows of k residues without an unmapped one ('X')
    bad = np.concatenate([[0], np.cumsum(res == ord("X"), dtype=np.int64)])
    lens = np.diff(off)
    starts_total = np.maximum(lens - (k - 1), 0)
    # windows containing an X: count per sequence = number of window starts s with bad[s+k]-bad[s] > 0
    win_bad = (bad[k:] - bad[:-k]) > 0                                  # for every start position of the buffer
    cs = np.concatenate([[0], np.cumsum(win_bad, dtype=np.int64)])
    lo = off[:-1]
    hi = lo + starts_total
    invalid = cs[hi] - cs[lo]
    want_rows = starts_total - invalid
    rows = C.sum(dim=1, dtype=torch.int64).cpu().numpy()
    assert np.array_equal(rows, want_rows)
    assert torch.equal(C.sum(dim=0, dtype=torch.int64), basis.counts)
    assert int(basis.counts.sum().item()) == int(want_rows.sum())
    # first-occurrence order: the first sequence's windows come first, in order of appearance
    first_seq = bytes(res[off[0]:off[1]]).decode()
    from snekmer_b200 import alphabet as A
    lut = np.frombuffer(A.lut(a), dtype=np.uint8)
    seen = []
    for i in range(len(first_seq) - k + 1):
        d = lut[np.frombuffer(first_seq[i:i + k].encode(), dtype=np.uint8)]
        if (d == 0xFF).any():
            continue
        c = int(d[0]) * 100 + int(d[1]) * 10 + int(d[2])
        if c not in seen:
            seen.append(c)
    assert codes[:len(seen)].tolist() == seen
    # uint16 transport == int32, determinism
    C16 = E.count_dense(batch, a, k, basis, dtype=torch.uint16)
    assert torch.equal(C16.to(torch.int32), C)
    assert torch.equal(E.count_dense(batch, a, k, basis), C)
    # CSR path: same non-zeros
    rowptr, cols, vals = E.count_csr(batch, a, k, basis)
    assert int(vals.sum().item()) == int(want_rows.sum())
    nnz_rows = (C != 0).sum(dim=1)
    assert torch.equal(rowptr[1:] - rowptr[:-1], nnz_rows.to(torch.int64))
    sample = torch.arange(0, n, 997, device=C.device)
    for s in sample[:200].tolist():
        r0, r1 = int(rowptr[s].item()), int(rowptr[s + 1].item())
        dense_row = torch.zeros(1000, dtype=torch.int32, device=C.device)
        dense_row[cols[r0:r1].long()] = vals[r0:r1]
        assert torch.equal(dense_row, C[s])
