"""BASELINE config C1 on the GPU path: the reference's own bundled learn/apply case (alphabet 2 = solvacc, k = 8,
.test/input_learnapp: 7,069 proteins, 40 synthetic annotations) through the rule bodies
vectorize -> learn -> merge -> eval_apply -> evaluate -> apply, compared with the files and matrices the UNMODIFIED
reference wrote (tests/golden/c1/, generator tests/golden/make_golden_c1.py).  Also merge_rule(base_counts=...)."""
import hashlib
import os

import numpy as np
import pandas as pd
import pytest

from oracle import skm_oracle as O
from util import GOLDEN, unpack_vecs

pytestmark = pytest.mark.gpu

C1 = os.path.join(GOLDEN, "c1")
FILES = ["UP000322080_2603819", "UP000322981_424902"]
A, K = 2, 8


def test_c1_rule_chain_reproduces_reference(tmp_path):
    from snekmer_b200 import rules as R

    d = np.load(os.path.join(C1, "c1_golden.npz"))
    ann = [os.path.join(C1, "c1.ann")]
    counts = []
    for nb in FILES:
        npz = str(tmp_path / f"{nb}.npz")
        R.vectorize_rule(os.path.join(C1, nb + ".fasta.gz"), npz, str(tmp_path / f"{nb}.kmers"), A, K)
        z = np.load(npz)
        assert list(z["kmerlist"]) == list(d[f"{nb}_kmerlist"]) and z["kmerlist"].dtype == d[f"{nb}_kmerlist"].dtype
        assert list(z["ids"]) == list(d[f"{nb}_ids"]) and list(z["lengths"]) == list(d[f"{nb}_lengths"])
        assert hashlib.sha256("\n".join(map(str, z["seqs"])).encode()).hexdigest() == str(d[f"{nb}_seqs_sha256"])
        assert z["vecs"].dtype == np.float64 and np.array_equal(z["vecs"].astype(np.uint8), unpack_vecs(d, f"{nb}_"))
        out = str(tmp_path / f"kmer-counts-{nb}.csv")
        R.learn_rule(npz, ann, out)
        assert open(out, "rb").read() == d[f"{nb}_counts_csv"].tobytes()            # byte-identical learn output
        counts.append(out)
    tot = str(tmp_path / "kmer-counts-total.csv")
    R.merge_rule(counts, tot)
    assert open(tot, "rb").read() == d["totals_csv"].tobytes()
    totA = str(tmp_path / "kmer-counts-totalA.csv")
    R.merge_rule(counts[:1], totA)
    assert open(totA, "rb").read() == d["totalsA_csv"].tobytes()
    # eval_apply: full matrices, then the top-2 masked files the evaluate rule reads
    score_files = []
    for nb in FILES:
        out = str(tmp_path / f"seq-annotation-scores-{nb}.csv")
        r = R.eval_apply_rule(str(tmp_path / f"{nb}.npz"), ann, tot, out, save_associations=True)
        ref = d[f"{nb}_eval_scores"]
        assert r.rows == list(d[f"{nb}_eval_rows"]) and r.annotations == list(d[f"{nb}_eval_cols"])
        assert np.max(np.abs(r.scores - ref)) < 1e-12
        i1, i2, s1, s2 = O.top2(ref)
        clear = (s1 - s2) > 1e-9
        assert np.array_equal(r.top1[clear], i1[clear])
        assert np.allclose(r.score1, s1, rtol=1e-5, atol=0) and np.allclose(r.score2, s2, rtol=1e-5, atol=1e-300)
        R.eval_apply_rule(str(tmp_path / f"{nb}.npz"), ann, tot, out, save_associations=False)
        score_files.append(out)
    conf, glob = str(tmp_path / "confidence-matrix.csv"), str(tmp_path / "global-confidence-scores.csv")
    R.evaluate_rule(score_files, conf, glob, modifier=20)
    got = pd.read_csv(glob)
    want = pd.read_csv(pd.io.common.BytesIO(d["global_confidence_csv"].tobytes()))
    assert list(got.columns) == list(want.columns) == ["Difference", "confidence", "weight", "sum"]
    # the Difference bins come from scores that agree to 1e-12, not bit for bit: a pair sitting on a rounding boundary
    # of round(., 2) may change bin, so the histograms are compared with a tolerance of a few counts and reported
    moved = int(np.abs(got["sum"].values - want["sum"].values).sum())
    assert moved <= 8, moved
    assert np.array_equal(got["weight"].values, want["weight"].values)
    assert np.nanmax(np.abs(got["confidence"].values - want["confidence"].values)) < 5e-3
    if moved == 0:
        assert open(glob, "rb").read() == d["global_confidence_csv"].tobytes()
        assert open(conf, "rb").read() == d["confidence_matrix_csv"].tobytes()
    # apply: the second proteome against the matrix learned from the first, with the learned confidence table
    summ = str(tmp_path / f"kmer-summary-{FILES[1]}.csv")
    r = R.apply_rule(str(tmp_path / f"{FILES[1]}.npz"), totA, glob, summ, save_associations=True,
                     out_scores=str(tmp_path / "apply-scores.csv"))
    ref = d["apply_scores"]
    assert r.rows == list(d["apply_rows"]) and r.annotations == list(d["apply_cols"])
    assert np.max(np.abs(r.scores - ref)) < 1e-12
    i1, i2, s1, s2 = O.top2(ref)
    clear = (s1 - s2) > 1e-9
    assert np.array_equal(r.top1[clear], i1[clear])
    table = pd.read_csv(summ)
    assert list(table.columns) == ["index", "Prediction", "Score", "delta", "Confidence"] and len(table) == 3686
    assert [str(x) for x in table["Prediction"][clear]] == [str(d["apply_cols"][i]) for i in i1[clear]]
    assert np.allclose(table["Score"], s1, rtol=1e-5, atol=0)
    near = np.abs(((s1 - s2) * 100) % 1 - 0.5) < 1e-7
    assert np.array_equal(table["delta"].values[~near], np.round(s1 - s2, 2)[~near])


def test_c1_reference_evaluator_inputs_give_identical_confidence_files(tmp_path):
    """evaluate_rule on score files holding the REFERENCE's own C1 matrices (top-2 masked, written like the reference
    writes them) -> byte-identical confidence-matrix.csv and global-confidence-scores.csv."""
    import pyarrow as pa
    from pyarrow import csv as pacsv

    from snekmer_b200 import rules as R

    d = np.load(os.path.join(C1, "c1_golden.npz"))
    paths = []
    for nb in FILES:
        S = d[f"{nb}_eval_scores"]
        order = np.argsort(-S, axis=1, kind="stable")[:, :2]
        keep = np.zeros_like(S, dtype=bool)
        keep[np.arange(len(S))[:, None], order] = True
        c = {str(a): np.where(keep, S, np.nan)[:, j] for j, a in enumerate(d[f"{nb}_eval_cols"])}
        c["__index_level_0__"] = [str(x) for x in d[f"{nb}_eval_rows"]]
        p = str(tmp_path / f"scores-{nb}.csv")
        pacsv.write_csv(pa.table(c), p)
        paths.append(p)
    conf, glob = str(tmp_path / "conf.csv"), str(tmp_path / "glob.csv")
    R.evaluate_rule(paths, conf, glob, modifier=20)
    assert open(glob, "rb").read() == d["global_confidence_csv"].tobytes()
    assert open(conf, "rb").read() == d["confidence_matrix_csv"].tobytes()


@pytest.mark.parametrize("case", ["same", "other_alphabet", "other_k", "two_files_same"])
def test_merge_rule_base_counts_branch(case, tmp_path):
    """learn.smk:496-579: incremental learning — the merged counts are added to a base kmer-counts-total.csv when
    its k-mer columns use the same letters and length, and written alone otherwise.  Byte-identical files."""
    from snekmer_b200 import rules as R

    d = np.load(os.path.join(GOLDEN, "merge_base.npz"))
    cA, cB = tmp_path / "kmer-counts-synA.csv", tmp_path / "kmer-counts-synB.csv"
    cA.write_bytes(d["countsA_csv"].tobytes())
    cB.write_bytes(d["countsB_csv"].tobytes())
    base = tmp_path / "base.csv"
    base.write_bytes(d[("same" if case == "two_files_same" else case) + "_base_csv"].tobytes())
    out = tmp_path / "kmer-counts-total.csv"
    files = [str(cA), str(cB)] if case == "two_files_same" else [str(cB)]
    R.merge_rule(files, str(out), base_counts=str(base))
    assert out.read_bytes() == d[f"{case}_merged_csv"].tobytes()
