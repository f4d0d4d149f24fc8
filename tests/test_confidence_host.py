"""Host side of the confidence evaluation (no GPU): the row classification the reference does with substring tests
(learn.smk:997-1008) and the curve arithmetic, against the oracle on the reference-written fixture."""
import os

import numpy as np

from oracle import skm_evaluator as OEV
from snekmer_b200 import confidence as CF
from util import GOLDEN


def _tricky():
    d = np.load(os.path.join(GOLDEN, "eval_confidence.npz"))
    return OEV.read_scores_csv(bytes(d["tricky_csv"]).decode())


def test_classify_rows_matches_oracle_with_and_without_truth_shortcut():
    S, labels, cols = _tricky()
    ht, hf, pred, bins, tf, known, diff = OEV.file_tables(S, labels, cols)
    want = np.where(~known, 0, np.where(tf, 1, 2)).astype(np.uint8)
    assert np.array_equal(CF.classify_rows(pred, labels, cols), want)
    # labels of the form <annotation>_known_<i> / <acc>_unknown_<i>: the truth index short-cuts rows predicted right
    truth = np.array([cols.index(lab.rsplit("_known_", 1)[0]) if "_known_" in lab and lab.rsplit("_known_", 1)[0] in cols else -1
                      for lab in labels])
    assert np.array_equal(CF.classify_rows(pred, labels, cols, truth), want)
    # the quirk itself: a prediction that is a substring of the label counts as True
    assert CF.classify_rows(np.array([0]), ["TIGR00012_known_3"], ["TIGR0001", "TIGR00012"]).tolist() == [1]
    assert CF.classify_rows(np.array([1]), ["TIGR0001_known_3"], ["TIGR0001", "TIGR00012"]).tolist() == [2]
    assert CF.classify_rows(np.array([0]), ["X_unknown_3"], ["X", "Y"]).tolist() == [0]


def test_accumulator_curve_arithmetic_matches_oracle():
    """ConfidenceAccumulator.add_histograms + finalize (host float64) == the oracle's evaluate, incl. a prior merge."""
    S, labels, cols = _tricky()
    ht, hf, *_ = OEV.file_tables(S, labels, cols)
    acc = CF.ConfidenceAccumulator()
    acc._files = 2                               # two files -> float columns, like the reference's running tables
    acc.add_histograms(ht, hf, cols)
    acc.add_histograms(ht, hf, cols)
    got = acc.finalize()
    want = OEV.evaluate([(S, labels, cols), (S, labels, cols)])
    assert got.rows == want["rows"] and np.array_equal(got.ratio, want["ratio"], equal_nan=True)
    assert np.array_equal(got.confidence, want["confidence"], equal_nan=True)
    assert np.array_equal(got.weight, want["weight"]) and np.array_equal(got.sum, want["sum"])
    prior = dict(confidence=np.linspace(0.2, 1.0, 101), weight=np.full(101, 50.0), sum=np.arange(101.0))
    got2 = acc.finalize(dict(prior, weight_is_int=False, sum_is_int=False), 0.5)
    want2 = OEV.evaluate([(S, labels, cols), (S, labels, cols)], prior, 0.5)
    assert np.array_equal(got2.confidence, want2["confidence"], equal_nan=True)
    assert np.array_equal(got2.weight, want2["weight"]) and np.array_equal(got2.sum, want2["sum"])
