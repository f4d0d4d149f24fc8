"""BASELINE config C1 — the reference's own bundled case (.test/config_learnapp.yaml: alphabet 2, k 8,
.test/input_learnapp, 7,069 proteins) — pins the CPU oracle on what the unmodified reference produced
(tests/golden/make_golden_c1.py -> tests/golden/c1/).  CPU only; also the merge rule's base-counts branch."""
import gzip
import hashlib
import io
import os

import numpy as np
import pandas as pd

from oracle import skm_oracle as O
from util import GOLDEN, csv_frame, read_ann, unpack_vecs

C1 = os.path.join(GOLDEN, "c1")
FILES = ["UP000322080_2603819", "UP000322981_424902"]
A, K = 2, 8


def read_fasta_gz(path):
    ids, seqs, cur = [], [], None
    with gzip.open(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if cur is not None:
                    seqs.append("".join(cur))
                parts = line[1:].split(None, 1)
                ids.append(parts[0] if parts else "")
                cur = []
            elif cur is not None:
                cur.append(line.strip())
    if cur is not None:
        seqs.append("".join(cur))
    return ids, seqs


def c1_oracle_state():
    """Everything the oracle derives for C1 (shared with the GPU test through import)."""
    d = np.load(os.path.join(C1, "c1_golden.npz"))
    ann = read_ann(os.path.join(C1, "c1.ann"))
    lut, syms = O.build_lut(A)
    per = {}
    for nb in FILES:
        ids, seqs = read_fasta_gz(os.path.join(C1, nb + ".fasta.gz"))
        res, offs = O.pack(seqs)
        si, pos, code, valid = O.window_codes(res, offs, lut, len(syms), K)
        basis, tot = O.basis_codes(si, pos, code, valid, 0)
        C = O.count_matrix(si, code, valid, len(seqs), basis)
        anns, M, nseq, totals, total_seqs = O.learn_matrix(ids, C, ann)
        per[nb] = dict(ids=ids, seqs=seqs, basis=basis, kmers=list(O.decode(basis, syms, K)), C=C, anns=anns, M=M, nseq=nseq,
                       totals=totals, total_seqs=total_seqs)
    return d, per


def test_c1_oracle_matches_reference_outputs():
    d, per = c1_oracle_state()
    for nb in FILES:
        p = per[nb]
        assert p["ids"] == list(d[f"{nb}_ids"])
        assert [len(s) for s in p["seqs"]] == list(d[f"{nb}_lengths"])
        red = [O.reduce_str(s, A) for s in p["seqs"]]
        assert hashlib.sha256("\n".join(red).encode()).hexdigest() == str(d[f"{nb}_seqs_sha256"])
        assert p["kmers"] == list(d[f"{nb}_kmerlist"]) and len(p["kmers"]) == 3 ** 8        # saturated space
        assert np.array_equal(O.presence_matrix(p["C"]).astype(np.uint8), unpack_vecs(d, f"{nb}_"))
        ref = csv_frame(d[f"{nb}_counts_csv"])
        assert list(ref.index) == ["Totals"] + p["anns"] and len(p["anns"]) == 40
        assert list(ref.columns) == ["Sequence count", "Kmer Count"] + p["kmers"]
        got = np.zeros((len(p["anns"]) + 1, len(p["kmers"]) + 2), dtype=np.int64)
        got[0, 0], got[0, 1], got[0, 2:] = p["total_seqs"], p["totals"].sum(), p["totals"]
        got[1:, 0], got[1:, 1], got[1:, 2:] = p["nseq"], p["M"].sum(axis=1), p["M"]
        assert np.array_equal(ref.values.astype(np.int64), got)
    # merged totals: outer join on k-mer columns, rows in order of first appearance
    tot = csv_frame(d["totals_csv"])
    a, b = per[FILES[0]], per[FILES[1]]
    seen = set(a["kmers"])
    cols = a["kmers"] + [x for x in b["kmers"] if x not in seen]
    rows = a["anns"] + [x for x in b["anns"] if x not in set(a["anns"])]
    assert list(tot.columns) == ["Sequence count", "Kmer Count"] + cols and list(tot.index) == ["Totals"] + rows
    cpos = {x: i for i, x in enumerate(cols)}
    Mm = np.zeros((len(rows), len(cols)), dtype=np.int64)
    for p in (a, b):
        ci = np.array([cpos[x] for x in p["kmers"]])
        for r, an in enumerate(p["anns"]):
            Mm[rows.index(an), ci] += p["M"][r]
    assert np.array_equal(tot.values[1:, 2:].astype(np.int64), Mm)
    assert int(tot.values[0, 0]) == a["total_seqs"] + b["total_seqs"] == 7069
    # eval_apply cosine of every sequence against the merged matrix
    for p, nb in ((a, FILES[0]), (b, FILES[1])):
        ci = np.array([cpos[x] for x in p["kmers"]])
        Q = np.zeros((p["C"].shape[0], len(cols)), dtype=np.int64)
        Q[:, ci] = p["C"]
        S = O.cosine_scores(Q, Mm)
        ref = d[f"{nb}_eval_scores"]
        assert list(d[f"{nb}_eval_cols"]) == rows and S.shape == ref.shape
        assert np.max(np.abs(S - ref)) < 1e-12
        i1, i2, s1, s2 = O.top2(S)
        r1 = np.argsort(-ref, axis=1, kind="stable")[:, 0]
        clear = (s1 - s2) > 1e-9
        assert np.array_equal(i1[clear], r1[clear])
    # apply: file 2 against the matrix learned from file 1
    assert list(d["apply_cols"]) == a["anns"] and list(d["apply_rows"]) == b["ids"]
    posA = {x: i for i, x in enumerate(a["kmers"])}
    Qs = np.zeros((b["C"].shape[0], len(a["kmers"])), dtype=np.int64)
    src = np.array([j for j, x in enumerate(b["kmers"]) if x in posA])
    dst = np.array([posA[b["kmers"][j]] for j in src])
    Qs[:, dst] = b["C"][:, src]
    S = O.cosine_scores(Qs, a["M"], q_norm_sq=(b["C"].astype(np.float64) ** 2).sum(axis=1))
    assert np.max(np.abs(S - d["apply_scores"])) < 1e-12


def test_c1_evaluator_oracle_matches_reference_confidence_files():
    """The oracle's confidence evaluation on the reference's own C1 score matrices == the files the reference's
    Evaluator wrote (bit-exact values and labels), conf_weight_modifier = 20 as in config_learnapp.yaml."""
    from oracle import skm_evaluator as EV

    d = np.load(os.path.join(C1, "c1_golden.npz"))
    inputs = []
    for nb in FILES:
        S = d[f"{nb}_eval_scores"]
        order = np.argsort(-S, axis=1, kind="stable")[:, :2]
        keep = np.zeros_like(S, dtype=bool)
        keep[np.arange(len(S))[:, None], order] = True
        inputs.append((np.where(keep, S, np.nan), [str(x) for x in d[f"{nb}_eval_rows"]], [str(x) for x in d[f"{nb}_eval_cols"]]))
    # the reference reads the matrices back from CSV text: go through the same text
    import pyarrow as pa
    from pyarrow import csv as pacsv
    parsed = []
    for S, rows, cols in inputs:
        c = {a: S[:, j] for j, a in enumerate(cols)}
        c["__index_level_0__"] = rows
        buf = io.BytesIO()
        pacsv.write_csv(pa.table(c), buf)
        parsed.append(EV.read_scores_csv(buf.getvalue().decode()))
    got = EV.evaluate(parsed, None, 20)
    import csv
    r = list(csv.reader(io.StringIO(d["global_confidence_csv"].tobytes().decode())))
    assert r[0] == ["Difference", "confidence", "weight", "sum"] and len(r) == 102
    want = np.array([[float(v) if v != "" else np.nan for v in x] for x in r[1:]])      # repr round trip: bit-exact
    assert np.array_equal(want[:, 1], got["confidence"], equal_nan=True)
    assert np.array_equal(want[:, 2], got["weight"]) and np.array_equal(want[:, 3], got["sum"])
    rc = list(csv.reader(io.StringIO(d["confidence_matrix_csv"].tobytes().decode())))
    assert [x[-1] for x in rc[1:]] == got["rows"]
    ratio = np.array([[float(v) if v != "" else np.nan for v in x[:-1]] for x in rc[1:]])
    assert np.array_equal(ratio, got["ratio"], equal_nan=True)


def test_merge_base_counts_oracle():
    """learn.smk:496-579: kmer-counts-total.csv merged with a base file (same letters / k) or not (different)."""
    d = np.load(os.path.join(GOLDEN, "merge_base.npz"))
    base, new = csv_frame(d["same_base_csv"]), csv_frame(d["countsB_csv"])
    merged = csv_frame(d["same_merged_csv"])
    # outer join: base rows / columns first, then the new ones; integer sums
    cols = list(base.columns) + [c for c in new.columns if c not in set(base.columns)]
    rows = list(base.index) + [r for r in new.index if r not in set(base.index)]
    assert list(merged.columns) == cols and list(merged.index) == rows
    want = (base.reindex(index=rows, columns=cols).fillna(0) + new.reindex(index=rows, columns=cols).fillna(0)).values.astype(np.int64)
    assert np.array_equal(merged.values.astype(np.int64), want)
    for other in ("other_alphabet", "other_k"):            # not merged: the output is the new counts alone
        out = csv_frame(d[f"{other}_merged_csv"])
        assert list(out.columns) == list(new.columns) and list(out.index) == list(new.index)
        assert np.array_equal(out.values.astype(np.int64), new.values.astype(np.int64))
