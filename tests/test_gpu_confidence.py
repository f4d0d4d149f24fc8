"""GPU parity of the confidence evaluation (learn.smk:923-1348): byte-identical output files against the golden
files written by the unmodified reference (tests/golden/eval_confidence.npz), the oracle on random inputs, and
the fused flow (top-2 straight from the scoring kernel, no Q x A matrix)."""
import csv
import io
import os

import numpy as np
import pytest
import torch

from oracle import skm_evaluator as OEV
from oracle import skm_oracle as O
from util import GOLDEN

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from snekmer_b200 import confidence as CF
    from snekmer_b200 import engine as E

D = np.load(os.path.join(GOLDEN, "eval_confidence.npz"))
CASES = {"one": (["synA"], None, 1.0), "two": (["synA", "synB"], None, 1.0), "tricky": (["tricky"], None, 1.0),
         "tricky_first": (["tricky", "synA"], None, 1.0), "prior": (["synB"], "one", 0.5)}


@pytest.mark.parametrize("name", list(CASES))
def test_evaluator_files_byte_identical(name, tmp_path):
    files, prior, mod = CASES[name]
    paths = []
    for f in files:
        p = tmp_path / f"seq-annotation-scores-{f}.csv"
        p.write_bytes(bytes(D[f"{f}_csv"]))
        paths.append(str(p))
    base = []
    if prior:
        bp = tmp_path / "prior.csv"
        bp.write_bytes(bytes(D[f"{prior}_glob"]))
        base = [str(bp)]
    ev = CF.Evaluator(paths, str(tmp_path / "conf.csv"), str(tmp_path / "glob.csv"), base, modifier=mod)
    ev.execute_all()
    assert (tmp_path / "glob.csv").read_bytes() == bytes(D[f"{name}_glob"])
    assert (tmp_path / "conf.csv").read_bytes() == bytes(D[f"{name}_conf"])


def test_top2_rows_and_bins_random():
    rng = np.random.default_rng(9)
    q, a = 3000, 37
    S = np.round(rng.random((q, a)), 2)
    S[rng.random((q, a)) < 0.3] = np.nan
    S[5] = np.nan
    S[6, :] = np.nan
    S[6, 3] = 0.5
    S[7] = 0.25
    r = CF.top2_rows(torch.from_numpy(S).cuda())
    pred, top, second = OEV.top_two_values(S)
    assert np.array_equal(r.top1.cpu().numpy(), pred)
    assert np.array_equal(r.score1.cpu().numpy(), top, equal_nan=True)
    assert np.array_equal(r.score2.cpu().numpy(), second, equal_nan=True)
    bins = CF.difference_bins(r.top1, r.score1, r.score2, a).cpu().numpy()
    diff = -(np.round(second - top, 2))
    want = np.where(np.isnan(diff), 255, np.rint(diff * 100)).astype(np.int64)
    assert np.array_equal(bins.astype(np.int64), want)
    # every bin value is one of the reference's 101 labels
    ok = bins != 255
    assert np.array_equal(np.array(CF.POSSIBLE_VALS)[bins[ok]], np.abs(diff[ok]))


def test_fused_flow_matches_oracle_on_scores():
    """learn -> apply -> confidence without the Q x A matrix == the oracle evaluating the full matrix."""
    rng = np.random.default_rng(4)
    aa = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
    fams = ["".join(rng.choice(aa, size=200)) for _ in range(12)]
    seqs, truth = [], []
    for i in range(900):
        f = int(rng.integers(0, 12))
        s = np.array(list(fams[f]))
        m = rng.random(len(s)) < 0.25
        s[m] = rng.choice(aa, size=int(m.sum()))
        seqs.append("".join(s))
        truth.append(f if rng.random() < 0.8 else -1)
    truth = np.array(truth)
    names = [f"FAM{f:02d}" for f in range(12)]
    a, k = 2, 5
    batch = E.SequenceBatch.from_strings(seqs)
    basis = E.build_basis(batch, a, k, 0)
    M, _ = E.learn_dense(batch, a, k, basis, torch.from_numpy(truth.astype(np.int32)), 12)
    Q = E.count_dense(batch, a, k, basis)
    labels = [f"{names[t]}_known_{i}" if t >= 0 else f"ACC{i}_unknown_{i}" for i, t in enumerate(truth)]
    half = 450
    acc = CF.ConfidenceAccumulator()
    for lo, hi in ((0, half), (half, 900)):
        r = E.apply_dense(Q[lo:hi], M[:12].contiguous())
        acc.add(r, labels[lo:hi], names, truth[lo:hi])
    got = acc.finalize()
    S = O.cosine_scores(Q.cpu().numpy(), M[:12].cpu().numpy())
    want = OEV.evaluate([(S[:half], labels[:half], names), (S[half:], labels[half:], names)])
    assert got.rows == want["rows"]
    assert np.array_equal(got.ratio, want["ratio"], equal_nan=True)
    assert np.array_equal(got.confidence, want["confidence"], equal_nan=True)
    assert np.array_equal(got.weight, want["weight"]) and np.array_equal(got.sum, want["sum"])
    assert got.zero_negative == want["zero_negative"]


def test_evaluate_results_equals_evaluate_rule(tmp_path):
    """eval_apply -> evaluate through files == through in-memory top-2 results (golden fixtures, miqs k=3)."""
    from snekmer_b200 import rules as R

    paths, results = [], []
    for f in ("synA", "synB"):
        p = tmp_path / f"seq-annotation-scores-{f}.csv"
        p.write_bytes(bytes(D[f"{f}_csv"]))
        paths.append(str(p))
        r = list(csv.reader(io.StringIO(bytes(D[f"{f}_csv"]).decode())))
        ix = r[0].index("__index_level_0__")
        cols = [c for i, c in enumerate(r[0]) if i != ix]
        S = np.array([[float(v) if v != "" else np.nan for i, v in enumerate(x) if i != ix] for x in r[1:]])
        t = CF.top2_rows(torch.from_numpy(S).cuda())
        results.append(R.ScoreResult([x[ix] for x in r[1:]], cols, t.top1.cpu().numpy(), t.top2.cpu().numpy(),
                                     t.score1.cpu().numpy(), t.score2.cpu().numpy(), None))
    R.evaluate_rule(paths, str(tmp_path / "c1.csv"), str(tmp_path / "g1.csv"))
    R.evaluate_results(results, str(tmp_path / "c2.csv"), str(tmp_path / "g2.csv"))
    assert (tmp_path / "g1.csv").read_bytes() == (tmp_path / "g2.csv").read_bytes() == bytes(D["two_glob"])
    assert (tmp_path / "c1.csv").read_bytes() == (tmp_path / "c2.csv").read_bytes() == bytes(D["two_conf"])


def test_delta_and_confidence_vectorised_equals_scalar():
    """apply.smk:312-335 for all queries at once (device bins + table look-up) == np.round + dict.get per row,
    including rows without a runner-up (one annotation) and a confidence file with missing keys."""
    from snekmer_b200 import rules as R

    rng = np.random.default_rng(12)
    q = 5000
    s1 = np.round(rng.random(q), 4)
    s2 = np.round(s1 * rng.random(q), 4)
    s2[::7] = s1[::7]                                   # exact ties
    s2[::11] = np.nan                                   # no runner-up
    conf = {v: 0.5 + 0.004 * i for i, v in enumerate(CF.POSSIBLE_VALS) if i % 13 != 5}       # some keys missing
    r = E.ApplyResult(torch.zeros(q, dtype=torch.int32, device="cuda"), torch.ones(q, dtype=torch.int32, device="cuda"),
                      torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda())
    delta, confidence = R.delta_and_confidence(r, conf, 3)
    want_delta = np.round(s1 - s2, 2)
    want_conf = np.array([conf.get(float(d), np.nan) for d in want_delta])
    assert np.array_equal(delta, want_delta, equal_nan=True)
    assert np.array_equal(confidence, want_conf, equal_nan=True)
