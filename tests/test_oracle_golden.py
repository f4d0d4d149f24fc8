"""Pin the CPU oracle (oracle/skm_oracle.py) on golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import skm_oracle as O
from util import GOLDEN, RULE_CONFIGS, csv_frame, edge_cases, load_rule, read_ann, read_fasta, unpack_vecs


def _alpha(a):
    return None if a == "None" else (int(a) if str(a).isdigit() else a)


def test_edge_reduce_vectorize_all_alphabets():
    fx = edge_cases()
    seqs = [s for _, s in fx["sequences"]]
    assert len(fx["cases"]) == 8 * 6
    for key, case in fx["cases"].items():
        a, k = key.split(":")
        a, k = _alpha(a), int(k)
        assert sorted(O.symbols_of(a)) == case["char_set"]
        lut, syms = O.build_lut(a)
        res, offs = O.pack(seqs)
        fits = len(syms) ** k <= 2 ** 64          # ptm (30 symbols) k=14 does not
        if fits:
            si, pos, code, valid = O.window_codes(res, offs, lut, len(syms), k)
        else:
            with pytest.raises(ValueError):
                O.window_codes(res, offs, lut, len(syms), k)
        for i, s in enumerate(seqs):
            assert O.reduce_str(s, a) == case["reduced"][i], (key, i)
            assert O.reduce_vectorize_str(s, a, k) == case["kmers"][i], (key, i)
            if not fits:
                continue
            sel = (si == i) & valid
            got = list(O.decode(code[sel], syms, k))
            assert got == case["kmers"][i], (key, i)


@pytest.mark.parametrize("name", sorted(RULE_CONFIGS))
def test_rule_level_fixtures(name):
    a, k, mf = RULE_CONFIGS[name]
    a = _alpha(a)
    d = load_rule(name)
    ann = read_ann(os.path.join(GOLDEN, "syn.ann"))
    lut, syms = O.build_lut(a)
    per_file = {}
    for nb in ("synA", "synB"):
        ids, seqs = read_fasta(os.path.join(GOLDEN, f"{nb}.fasta"))
        assert ids == list(d[f"{nb}_ids"])
        assert [len(s) for s in seqs] == list(d[f"{nb}_lengths"])
        assert [O.reduce_str(s, a) for s in seqs] == list(d[f"{nb}_seqs"])
        res, offs = O.pack(seqs)
        si, pos, code, valid = O.window_codes(res, offs, lut, len(syms), k)
        basis, tot = O.basis_codes(si, pos, code, valid, mf)
        assert list(O.decode(basis, syms, k)) == list(d[f"{nb}_kmerlist"])
        assert O.basis_str(seqs, a, k, mf) == list(d[f"{nb}_kmerlist"])
        C = O.count_matrix(si, code, valid, len(seqs), basis)
        assert np.array_equal(O.presence_matrix(C), unpack_vecs(d, f"{nb}_").astype(np.float64))
        # learn counts CSV (learn.smk:328-357)
        anns, M, nseq, totals, total_seqs = O.learn_matrix(ids, C, ann)
        ref = csv_frame(d[f"{nb}_counts_csv"])
        assert list(ref.index) == ["Totals"] + anns
        assert list(ref.columns) == ["Sequence count", "Kmer Count"] + list(d[f"{nb}_kmerlist"])
        got = np.zeros((len(anns) + 1, len(basis) + 2), dtype=np.int64)
        got[0, 0], got[0, 1], got[0, 2:] = total_seqs, totals.sum(), totals
        got[1:, 0], got[1:, 1], got[1:, 2:] = nseq, M.sum(axis=1), M
        assert np.array_equal(ref.values.astype(np.int64), got)
        # string-level counts agree with the integer formulation
        for i in (0, len(seqs) // 2, len(seqs) - 1):
            assert O.counts_str(seqs[i], list(d[f"{nb}_kmerlist"]), a) == list(C[i])
        per_file[nb] = (ids, basis, C, anns, M, totals)

    # merged totals (learn.smk:467-494): outer join on k-mer, row order of first appearance
    tot = csv_frame(d["totals_csv"])
    kmA = list(O.decode(per_file["synA"][1], syms, k))
    kmB = list(O.decode(per_file["synB"][1], syms, k))
    cols = kmA + [x for x in kmB if x not in set(kmA)]
    assert list(tot.columns) == ["Sequence count", "Kmer Count"] + cols
    rows = list(per_file["synA"][3]) + [x for x in per_file["synB"][3] if x not in per_file["synA"][3]]
    assert list(tot.index) == ["Totals"] + rows
    Mm = np.zeros((len(rows), len(cols)), dtype=np.int64)
    for nb, km in (("synA", kmA), ("synB", kmB)):
        _, _, _, anns, M, _ = per_file[nb]
        ci = [cols.index(x) for x in km]
        for r, an in enumerate(anns):
            Mm[rows.index(an), ci] += M[r]
    assert np.array_equal(tot.values[1:, 2:].astype(np.int64), Mm)

    # eval_apply cosine (learn.smk:811-829): query columns = own basis, M columns = merged
    for nb, km in (("synA", kmA), ("synB", kmB)):
        ids, basis, C, *_ = per_file[nb]
        ci = np.array([cols.index(x) for x in km])
        Q = np.zeros((C.shape[0], len(cols)), dtype=np.int64)
        Q[:, ci] = C
        S = O.cosine_scores(Q, Mm)
        ref = d[f"{nb}_eval_scores"]
        assert list(d[f"{nb}_eval_cols"]) == rows
        assert S.shape == ref.shape
        assert np.max(np.abs(S - ref)) < 1e-12
        i1, i2, s1, s2 = O.top2(S)
        assert np.array_equal(i1, np.argsort(-ref, axis=1, kind="stable")[:, 0]) or \
            np.allclose(ref[np.arange(len(i1)), i1], ref.max(axis=1), rtol=0, atol=1e-12)

    # apply (apply.smk:224-289): synB queries vs matrix learned on synA only;
    # query norm spans the query's own basis, the dot only the shared k-mers
    _, basisB, CB, *_ = per_file["synB"]
    _, _, _, annsA, MA, _ = per_file["synA"]
    assert list(d["apply_cols"]) == annsA
    posA = {x: i for i, x in enumerate(kmA)}
    shared = [(j, posA[x]) for j, x in enumerate(kmB) if x in posA]
    Qs = np.zeros((CB.shape[0], len(kmA)), dtype=np.int64)
    for j, ia in shared:
        Qs[:, ia] = CB[:, j]
    S = O.cosine_scores(Qs, MA, q_norm_sq=(CB.astype(np.float64) ** 2).sum(axis=1))
    assert np.max(np.abs(S - d["apply_scores"])) < 1e-12
    assert list(d["apply_rows"]) == per_file["synB"][0]


def test_basis_txt_branch():
    d = np.load(os.path.join(GOLDEN, "basis_txt.npz"))
    ids, seqs = read_fasta(os.path.join(GOLDEN, "synA.fasta"))
    assert list(d["kmerlist"]) == list(d["basis"])
    lut, syms = O.build_lut(2)
    res, offs = O.pack(seqs)
    si, pos, code, valid = O.window_codes(res, offs, lut, len(syms), 3)
    ok = [all(c in syms for c in km) for km in d["basis"]]
    codes = O.encode_kmers([km for km, o in zip(d["basis"], ok) if o], syms)
    C = np.zeros((len(seqs), len(d["basis"])), dtype=np.int32)
    C[:, np.flatnonzero(ok)] = O.count_matrix(si, code, valid, len(seqs), codes)
    n = int(np.prod(d["vecs_shape"]))
    ref = np.unpackbits(d["vecs_bits"])[:n].reshape(tuple(d["vecs_shape"]))
    assert np.array_equal((C > 0).astype(np.uint8), ref)


def test_top2_ties_and_rounding():
    S = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.1], [0.1, 0.7, 0.7]])
    i1, i2, s1, s2 = O.top2(S)
    assert list(i1) == [0, 0, 1] and list(i2) == [1, 1, 2]
    assert np.round(0.125, 2) == 0.12 and np.round(0.135, 2) == 0.14
    pred, score, delta, conf = O.apply_table(S, ["a", "b", "c"], {0.0: 0.25})
    assert pred == ["a", "a", "b"] and list(delta) == [0.0, 0.0, 0.0] and list(conf) == [0.25] * 3


def test_csr_matches_dense():
    ids, seqs = read_fasta(os.path.join(GOLDEN, "synB.fasta"))
    lut, syms = O.build_lut("miqs")
    res, offs = O.pack(seqs)
    si, pos, code, valid = O.window_codes(res, offs, lut, len(syms), 3)
    rowptr, ccodes, ccnt = O.count_csr(si, code, valid, len(seqs))
    full = np.arange(len(syms) ** 3, dtype=np.uint64)
    C = O.count_matrix(si, code, valid, len(seqs), full)
    for r in range(len(seqs)):
        nz = np.flatnonzero(C[r])
        assert np.array_equal(nz.astype(np.uint64), ccodes[rowptr[r]:rowptr[r + 1]])
        assert np.array_equal(C[r, nz], ccnt[rowptr[r]:rowptr[r + 1]])


# ---------------------------------------------------------------------------
# confidence evaluation (learn.smk:923-1348): oracle/skm_evaluator.py pinned on files written by the reference
# ---------------------------------------------------------------------------
def _eval_fixture():
    import csv
    import io

    from oracle import skm_evaluator as EV

    d = np.load(os.path.join(GOLDEN, "eval_confidence.npz"))
    txt = lambda k: bytes(d[k]).decode()
    scores = EV.read_scores_csv

    def glob(t):
        r = list(csv.reader(io.StringIO(t)))
        assert r[0] == ["Difference", "confidence", "weight", "sum"]
        return [x[0] for x in r[1:]], np.array([[float(v) if v != "" else np.nan for v in x] for x in r[1:]])

    def conf(t):
        r = list(csv.reader(io.StringIO(t)))
        assert r[0][-1] == "Prediction"
        rows = [x[-1] for x in r[1:]]
        return r[0][:-1], rows, np.array([[float(v) if v != "" else np.nan for v in x[:-1]] for x in r[1:]]).reshape(len(rows), -1)

    return EV, txt, scores, glob, conf


def test_pandas_float_restatement_matches_pandas():
    """oracle.skm_evaluator.pandas_float == the float converter of pandas.read_csv (C engine, default precision),
    which is what the reference's Evaluator reads its inputs with — and it is NOT float(): ~30 % of 17-digit reprs
    come back one ulp off."""
    import io
    import random

    import pandas as pd

    from oracle import skm_evaluator as EV

    rng = random.Random(3)
    tests = [repr((rng.random() * 10 ** rng.randint(-8, 8)) * (-1 if rng.random() < 0.1 else 1)) for _ in range(40000)]
    tests += ["1e-05", "2.5e-07", "1.0", "0.0", "-0.0", "1", "0.5", "123456789012345678", "1e22", "3.0000000000000004e-05",
              "1e-300", "4.9e-324", "1.7976931348623157e308", "0.9099999999999999", "0.30000000000000004"]
    got = pd.read_csv(io.StringIO("a\n" + "\n".join(tests) + "\n"), dtype=float)["a"].values
    mine = np.array([EV.pandas_float(t) for t in tests])
    assert np.array_equal(mine, got)
    assert EV.pandas_float("0.9099999999999999") == 0.91 != float("0.9099999999999999")
    assert (got != np.array([float(t) for t in tests])).mean() > 0.1


@pytest.mark.parametrize("name,files,prior,mod", [("one", ["synA"], None, 1.0), ("two", ["synA", "synB"], None, 1.0),
                                                  ("tricky", ["tricky"], None, 1.0), ("tricky_first", ["tricky", "synA"], None, 1.0),
                                                  ("prior", ["synB"], "one", 0.5)])
def test_evaluator_oracle_matches_reference_files(name, files, prior, mod):
    EV, txt, scores, glob, conf = _eval_fixture()
    pr = EV.read_global_csv(txt(prior + "_glob")) if prior else None
    r = EV.evaluate([scores(txt(f + "_csv")) for f in files], pr, mod)
    lab, g = glob(txt(name + "_glob"))
    hdr, rows, c = conf(txt(name + "_conf"))
    assert len(lab) == 101 and lab[0] == ("-0.0" if r["zero_negative"] else "0.0") and hdr[0] == lab[0]
    assert [float(x) for x in lab[1:]] == list(EV.POSSIBLE[1:])
    assert np.array_equal(g[:, 1], r["confidence"], equal_nan=True)          # bit-exact float64
    assert np.array_equal(g[:, 2], r["weight"]) and np.array_equal(g[:, 3], r["sum"])
    assert rows == r["rows"] and np.array_equal(c, r["ratio"], equal_nan=True)
