"""Binary side-cars (snekmer_b200/sidecar.py): container round trips and lossless conversion to / from the
reference's own files — the CSV / npz bytes written by the unmodified reference (tests/golden/rule_*.npz)."""
import os

import numpy as np
import pytest

from snekmer_b200 import sidecar as SC
from snekmer_b200 import rules as R
from util import GOLDEN, load_rule


def test_container_roundtrip(tmp_path):
    arrays = {"a": np.arange(10, dtype=np.int64), "b": np.zeros((0, 3), dtype=np.float32), "c": np.frombuffer(b"xyz", dtype=np.uint8),
              "d": np.random.default_rng(0).random((5, 7))}
    p = str(tmp_path / "x.skm")
    SC.write_container(p, "test", {"k": 3, "name": "ünï"}, arrays)
    for mmap in (True, False):
        kind, attrs, got = SC.read_container(p, mmap)
        assert kind == "test" and attrs == {"k": 3, "name": "ünï"}
        for k, v in arrays.items():
            assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    with open(p, "r+b") as f:
        f.write(b"BAD")
    with pytest.raises(ValueError):
        SC.read_container(p)


@pytest.mark.parametrize("name", ["miqs_k3", "standard_k5_mf1", "hydro_k8"])
def test_counts_csv_roundtrip_bytes(name, tmp_path):
    """reference CSV -> .skmc -> CSV is byte-identical, for the per-file layout and the merged layout."""
    d = load_rule(name)
    for key, exporter in (("synA_counts_csv", SC.export_counts_csv), ("totals_csv", SC.export_totals_csv)):
        src = tmp_path / f"{key}.csv"
        src.write_bytes(bytes(d[key]))
        side = str(tmp_path / f"{key}.skmc")
        SC.import_counts_csv(str(src), side)
        out = tmp_path / f"{key}.out.csv"
        exporter(side, str(out))
        assert out.read_bytes() == src.read_bytes(), key
        t0, t1 = R.read_counts_csv(str(src)), SC.read_counts(side)
        assert t0.rows == t1.rows and t0.kmers == t1.kmers and np.array_equal(t0.M, t1.M)


def test_vectors_npz_export_matches_reference(tmp_path):
    """.skmv built from the reference's own vectors -> export_npz reproduces every array of the rule's .npz."""
    d = load_rule("solvacc_k4")
    kmerlist = [str(x) for x in d["synA_kmerlist"]]
    shape = tuple(d["synA_vecs_shape"])
    vecs = np.unpackbits(d["synA_vecs_bits"])[: shape[0] * shape[1]].reshape(shape)
    symbols = "".join(sorted(set("".join(kmerlist))))
    idx = {c: i for i, c in enumerate(symbols)}
    codes = np.array([sum(idx[ch] * len(symbols) ** (4 - 1 - j) for j, ch in enumerate(km)) for km in kmerlist], dtype=np.uint64)
    rowptr, cols, vals = SC.dense_to_csr(vecs)
    p = str(tmp_path / "a.skmv")
    SC.write_vectors(p, "solvacc", 4, symbols, codes, d["synA_ids"], d["synA_seqs"], d["synA_lengths"], rowptr, cols, vals)
    out = str(tmp_path / "a.npz")
    SC.export_npz(p, out)
    z = np.load(out)
    assert list(z["kmerlist"]) == kmerlist and list(z["ids"]) == list(d["synA_ids"]) and list(z["seqs"]) == list(d["synA_seqs"])
    assert np.array_equal(z["lengths"], d["synA_lengths"]) and z["vecs"].dtype == np.float64 and np.array_equal(z["vecs"], vecs.astype(np.float64))


def test_counts_csr_roundtrip_without_dense(tmp_path):
    rows = ["Totals", "famA", "famB"]
    kmers = ["AAA", "AAC", "ACA", "CAA"]
    rowptr = np.array([0, 3, 5, 6])
    cols = np.array([0, 1, 3, 0, 3, 1], dtype=np.int32)
    vals = np.array([7, 2, 5, 4, 5, 2], dtype=np.int64)
    p = str(tmp_path / "m.skmc")
    SC.write_counts_csr(p, rows, kmers, np.array([9, 4, 5]), np.array([14, 9, 2]), rowptr, cols, vals)
    d = SC.read_counts_csr(p)
    assert d["rows"] == rows and d["kmers"] == kmers
    assert np.array_equal(d["rowptr"], rowptr) and np.array_equal(d["cols"], cols) and np.array_equal(d["vals"], vals)
    t = SC.read_counts(p)                       # the dense view of the same file
    want = np.zeros((3, 4), dtype=np.int64)
    want[np.repeat(np.arange(3), np.diff(rowptr)), cols] = vals
    assert np.array_equal(t.M, want) and t.seq_count.tolist() == [9, 4, 5] and t.kmer_count.tolist() == [14, 9, 2]
