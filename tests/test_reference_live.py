"""Live parity of the oracle against the UNMODIFIED reference, where /root/reference exists (the build container;
skipped on the GPU box).  Complements the committed golden vectors with fresh random inputs."""
import csv
import io
import os

import numpy as np
import pytest

import refharness as rh
from oracle import skm_evaluator as OEV
from oracle import skm_oracle as O

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


def _rand_seqs(rng, n, lo=0, hi=120):
    aa = np.array(list("ACDEFGHIKLMNPQRSTVWYXBZUO*acd"))
    p = np.array([1.0] * 20 + [0.05] * 9)
    return ["".join(rng.choice(aa, size=int(rng.integers(lo, hi)), p=p / p.sum())) for _ in range(n)]


@pytest.mark.parametrize("a", [0, 1, 2, 3, 4, 5, "ptm", None])
def test_live_reduce_and_kmers(a):
    skm = rh.load_reference()
    rng = np.random.default_rng(17)
    seqs = _rand_seqs(rng, 60) + ["", "A", "*", "AC*", "*AC", "ACDEF***"]
    for k in (1, 3, 6):
        kv = skm.vectorize.KmerVec(alphabet=a, k=k)
        for s in seqs:
            assert skm.vectorize.reduce(s, alphabet=a, mapping=skm.alphabet.FULL_ALPHABETS) == O.reduce_str(s, a)
            assert list(map(str, kv.reduce_vectorize(s))) == O.reduce_vectorize_str(s, a, k)


def test_live_evaluator_random(tmp_path):
    """The reference's Evaluator (exec'd from learn.smk:923-1348) on random top-2-masked score files == the oracle,
    value for value (float64 bit patterns), including the prior merge."""
    import pyarrow as pa
    from pyarrow import csv as pacsv

    rng = np.random.default_rng(23)
    cols = [f"FAM{i:03d}" for i in range(9)] + ["FAM00", "known"]
    files, paths = [], []
    for f in range(3):
        q = 150 + 40 * f
        S = np.round(rng.random((q, len(cols))), 2)                 # 2 decimals: many exact ties and bin edges
        truth = rng.integers(0, len(cols), size=q)
        boost = rng.random(q) < 0.6
        S[np.arange(q)[boost], truth[boost]] = np.minimum(1.0, S[np.arange(q)[boost], truth[boost]] + 0.3)
        order = np.argsort(-S, axis=1, kind="stable")[:, :2]
        keep = np.zeros_like(S, dtype=bool)
        keep[np.arange(q)[:, None], order] = True
        S = np.where(keep, S, np.nan)
        rows = [f"{cols[t]}_known_{i}" if i % 4 else f"Q{i}_unknown_{i}" for i, t in enumerate(truth)]
        p = str(tmp_path / f"scores-{f}.csv")
        c = {a: S[:, j] for j, a in enumerate(cols)}
        c["__index_level_0__"] = rows
        pacsv.write_csv(pa.table(c), p)
        files.append(OEV.read_scores_csv(open(p).read()))          # what the reference's read_csv sees (1-ulp parser)
        paths.append(p)

    def glob(path):
        r = list(csv.reader(open(path)))
        return [x[0] for x in r[1:]], np.array([[float(v) if v != "" else np.nan for v in x] for x in r[1:]])

    def conf(path):
        r = list(csv.reader(open(path)))
        rows = [x[-1] for x in r[1:]]
        return rows, np.array([[float(v) if v != "" else np.nan for v in x[:-1]] for x in r[1:]]).reshape(len(rows), -1)

    c1, g1 = rh.run_evaluate(str(tmp_path), paths[:2], "c1.csv", "g1.csv")
    want = OEV.evaluate(files[:2])
    lab, g = glob(g1)
    rows, c = conf(c1)
    assert lab[0] == ("-0.0" if want["zero_negative"] else "0.0")
    assert np.array_equal(g[:, 1], want["confidence"], equal_nan=True)
    assert np.array_equal(g[:, 2], want["weight"]) and np.array_equal(g[:, 3], want["sum"])
    assert rows == want["rows"] and np.array_equal(c, want["ratio"], equal_nan=True)
    # third file with the first result as prior
    c2, g2 = rh.run_evaluate(str(tmp_path), paths[2:], "c2.csv", "g2.csv", [g1], 0.25)
    want2 = OEV.evaluate(files[2:], OEV.read_global_csv(open(g1).read()), 0.25)
    _, gg = glob(g2)
    assert np.array_equal(gg[:, 1], want2["confidence"], equal_nan=True)
    assert np.array_equal(gg[:, 2], want2["weight"]) and np.array_equal(gg[:, 3], want2["sum"])
