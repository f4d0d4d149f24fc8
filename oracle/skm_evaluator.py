"""CPU oracle of the confidence evaluation (rule ``evaluate``, class Evaluator, learn.smk:923-1348).

TEST INFRASTRUCTURE ONLY (see oracle/skm_oracle.py).  A numpy restatement, without pandas, of what the
reference computes from the ``seq-annotation-scores-*.csv`` files of eval_apply:

  per file (learn.smk:964-1037)   prediction = column of the row maximum (NaN skipped, first maximum);
                                  Top / Second = the two largest values; Difference = -(round(Second - Top, 2));
                                  T iff the predicted name is a SUBSTRING of the row label; Known iff the label
                                  does not contain "unknown";
  crosstabs (learn.smk:1039-1176) counts of Known rows per (Prediction, Difference), T and F separately,
                                  summed over files, on the 101 columns 0.00 .. 1.00, rows = the sorted names
                                  of predictions that occur;
  outputs (learn.smk:1195-1318)   confidence matrix T / (T + F); global curve = column sums T / (T + F),
                                  linearly interpolated over empty bins (leading bins stay empty, trailing
                                  bins repeat the last value); weight = number of Known rows, sum = rows per
                                  bin; optional merge with ONE prior global-confidence file.

Third-party arithmetic at the file boundary: the reference reads the score files and the prior confidence file with
``pandas.read_csv`` (C engine, default ``float_precision``), whose float converter is NOT round-trip exact — it
accumulates at most 17 digits in a double and divides once by a power of ten (pandas/_libs/src/parser/tokenizer.c,
``precise_xstrtod``); about a third of 17-digit reprs come back one ulp off (``"0.9099999999999999"`` -> 0.91), which
moves exact ties and bin edges.  ``pandas_float`` restates that converter (pinned against pandas 3.0.2 on 200 k
strings, tests/test_oracle_golden.py) and ``read_scores_csv`` / ``read_global_csv`` use it, so the oracle sees the
numbers the reference sees.

Pinned against the unmodified reference run in the build container (tests/golden/make_golden_eval.py ->
tests/golden/eval_*.npz; tests/test_oracle_golden.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NBINS = 101
POSSIBLE = np.array([round(x * 0.01, 2) for x in range(NBINS)], dtype=np.float64)   # learn.smk:1063


def pandas_float(txt: str) -> float:
    """``precise_xstrtod`` of pandas' C parser (the default float converter of read_csv): sign, at most 17 digits
    accumulated as ``number * 10 + digit`` in a double (further integer digits bump the exponent, further decimals
    are dropped), optional exponent, then ONE multiplication / division by the double 1e|exponent|."""
    p, n, neg = 0, len(txt), False
    if p < n and txt[p] in "+-":
        neg = txt[p] == "-"
        p += 1
    number, exponent, num_digits, num_decimals, max_digits = 0.0, 0, 0, 0, 17
    while p < n and txt[p].isdigit():
        if num_digits < max_digits:
            number = number * 10.0 + (ord(txt[p]) - 48)
            num_digits += 1
        else:
            exponent += 1
        p += 1
    if p < n and txt[p] == ".":
        p += 1
        while num_digits < max_digits and p < n and txt[p].isdigit():
            number = number * 10.0 + (ord(txt[p]) - 48)
            p += 1
            num_digits += 1
            num_decimals += 1
        while p < n and txt[p].isdigit():
            p += 1
        exponent -= num_decimals
    if neg:
        number = -number
    if p < n and txt[p] in "eE":
        p += 1
        eneg = False
        if p < n and txt[p] in "+-":
            eneg = txt[p] == "-"
            p += 1
        e = 0
        while p < n and txt[p].isdigit():
            e = e * 10 + (ord(txt[p]) - 48)
            p += 1
        exponent += -e if eneg else e
    if exponent > 308:
        return float("inf")
    if exponent > 0:
        return number * float("1e%d" % exponent)
    if exponent < -308:
        return 0.0 if exponent < -616 else number / float("1e%d" % (-308 - exponent)) / 1e308
    return number / float("1e%d" % (-exponent))


_NA = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN", "<NA>", "N/A", "NA", "NULL",
       "NaN", "None", "n/a", "nan", "null"}          # pandas' default na_values


def _cell(v: str) -> float:
    return np.nan if v in _NA else pandas_float(v)


def read_scores_csv(text: str):
    """A seq-annotation-scores CSV as the reference's read_csv sees it -> (S float64 [Q, A], row labels, column names)."""
    import csv
    import io

    r = list(csv.reader(io.StringIO(text)))
    ix = r[0].index("__index_level_0__")
    cols = [c for i, c in enumerate(r[0]) if i != ix]
    S = np.array([[_cell(v) for i, v in enumerate(x) if i != ix] for x in r[1:]], dtype=np.float64).reshape(len(r) - 1, len(cols))
    return S, [x[ix] for x in r[1:]], cols


def read_global_csv(text: str) -> Dict[str, np.ndarray]:
    """A global-confidence-scores CSV (prior) as the reference's read_csv sees it."""
    import csv
    import io

    r = list(csv.reader(io.StringIO(text)))
    a = np.array([[_cell(v) for v in x] for x in r[1:]], dtype=np.float64).reshape(len(r) - 1, len(r[0]))
    return {h: a[:, i] for i, h in enumerate(r[0])}


def top_two_values(S: np.ndarray):
    """learn.smk:964-981: idxmax over the columns (NaN skipped, first maximum) and the two largest values of
    every row (``argpartition(-values, [0, 1])``: NaN sorts last).  Rows without any value get prediction -1."""
    S = np.asarray(S, dtype=np.float64)
    q, a = S.shape
    neg = np.where(np.isnan(S), -np.inf, S)
    pred = np.argmax(neg, axis=1)                       # first maximum
    top = neg[np.arange(q), pred]
    masked = neg.copy()
    masked[np.arange(q), pred] = -np.inf
    second = masked.max(axis=1) if a > 1 else np.full(q, -np.inf)
    pred = np.where(np.isneginf(top), -1, pred)
    return pred, np.where(np.isneginf(top), np.nan, top), np.where(np.isneginf(second), np.nan, second)


def file_tables(S: np.ndarray, labels: Sequence[str], annotations: Sequence[str]):
    """One seq-annotation-scores file -> (hist_true, hist_false) int64 [A, 101] indexed by PREDICTION, and the
    per-row (prediction, bin, tf, known) lists.  Rows whose Difference is NaN or not one of the 101 values drop
    out (pd.crosstab drops NaN; learn.smk:1104-1110 keeps only the possible values)."""
    pred, top, second = top_two_values(S)
    diff = -(np.round(second - top, 2))
    a = len(annotations)
    ht = np.zeros((a, NBINS), dtype=np.int64)
    hf = np.zeros((a, NBINS), dtype=np.int64)
    bins = np.full(len(labels), -1, dtype=np.int64)
    tf, known = [], []
    for i, lab in enumerate(labels):
        p = int(pred[i])
        t = p >= 0 and (str(annotations[p]) in str(lab))            # learn.smk:1000: substring test
        kn = "unknown" not in str(lab)                              # learn.smk:1004
        tf.append(t)
        known.append(kn)
        d = diff[i]
        if p < 0 or np.isnan(d):
            continue
        j = int(np.rint(d * 100.0))
        if j < 0 or j >= NBINS or POSSIBLE[j] != d:
            continue
        bins[i] = j
        if kn:
            (ht if t else hf)[p, j] += 1
    return ht, hf, pred, bins, np.array(tf), np.array(known), diff


def zero_label_is_negative(diff: np.ndarray, tf: np.ndarray, known: np.ndarray) -> bool:
    """The label of the zero bin in the outputs is the first zero Difference among the Known & T rows of the
    FIRST file (pd.crosstab keeps the first-seen representative of {0.0, -0.0}, and every later step keeps the
    label of the running table): ``-(round(0.0))`` = -0.0 for an exact tie Top == Second, +0.0 otherwise."""
    sel = known & tf & (diff == 0)
    idx = np.flatnonzero(sel)
    return bool(idx.size and np.signbit(diff[idx[0]]))


def interpolate_linear(y: np.ndarray) -> np.ndarray:
    """``Series.interpolate(method="linear")`` (learn.smk:1238): equally spaced points; NaNs between two values
    are filled linearly, trailing NaNs repeat the last value, leading NaNs stay."""
    y = np.asarray(y, dtype=np.float64).copy()
    ok = np.flatnonzero(~np.isnan(y))
    if ok.size == 0:
        return y
    x = np.arange(len(y), dtype=np.float64)
    inside = (x >= ok[0])
    y[inside] = np.interp(x[inside], ok.astype(np.float64), y[ok])
    return y


def evaluate(files: Sequence[Tuple[np.ndarray, Sequence[str], Sequence[str]]], prior: Optional[Dict[str, np.ndarray]] = None,
             modifier: float = 1.0):
    """files: [(S [Q, A] float64 with NaN, row labels, column names), ...].  prior: the columns of a
    global-confidence-scores.csv ("confidence", "weight", "sum", each float64 [101]) or None.
    Returns dict(rows = sorted prediction names, ratio [R, 101] (NaN = empty), confidence / weight / sum [101],
    zero_negative)."""
    names: Dict[str, int] = {}
    acc_t: List[np.ndarray] = []
    acc_f: List[np.ndarray] = []
    zero_neg = False
    for n, (S, labels, annotations) in enumerate(files):
        ht, hf, pred, bins, tf, known, diff = file_tables(S, labels, annotations)
        if n == 0:
            zero_neg = zero_label_is_negative(diff, tf, known)
        for j, name in enumerate(annotations):
            if ht[j].any() or hf[j].any():
                i = names.setdefault(str(name), len(names))
                if i == len(acc_t):
                    acc_t.append(np.zeros(NBINS, np.int64))
                    acc_f.append(np.zeros(NBINS, np.int64))
                acc_t[i] += ht[j]
                acc_f[i] += hf[j]
    rows = sorted(names)
    T = np.array([acc_t[names[r]] for r in rows], dtype=np.float64).reshape(len(rows), NBINS)
    F = np.array([acc_f[names[r]] for r in rows], dtype=np.float64).reshape(len(rows), NBINS)
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = T / (T + F)                                                    # learn.smk:1203-1205
        tt, ff = T.sum(axis=0), F.sum(axis=0)
        conf = interpolate_linear(tt / (tt + ff))                              # learn.smk:1235-1238
    sum_series = tt + ff
    weight = float(T.sum() + F.sum())
    out_w = np.full(NBINS, weight)
    out_sum = sum_series
    if prior is not None:                                                      # learn.smk:1263-1291
        pw = np.asarray(prior["weight"], dtype=np.float64)
        k_factor = 1 + modifier * (weight / (weight + pw))
        out_w = pw + weight
        weighted_current = k_factor * weight
        total_weight = pw + weighted_current
        conf = (np.asarray(prior["confidence"], np.float64) * pw + conf * weighted_current) / total_weight
        out_sum = sum_series + np.asarray(prior["sum"], np.float64)
    return dict(rows=rows, ratio=ratio, confidence=conf, weight=out_w, sum=out_sum, zero_negative=zero_neg)
