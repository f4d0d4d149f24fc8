"""The drop-in Python API and the rewritten rule bodies against the golden files produced
by the unmodified reference (tests/golden/make_golden.py).

CPU part: host-only pieces (utils, io, reduce, KmerBasis argument checks).
GPU part: KmerVec, vectorize / learn / merge / eval_apply / apply rules — file outputs are
compared byte for byte where the reference's writer is deterministic (kmer-counts and
totals CSVs), value for value otherwise."""
import io
import os
import pickle

import numpy as np
import pandas as pd
import pytest
import torch

import snekmer_b200 as skm
from oracle import skm_oracle as O
from util import GOLDEN, RULE_CONFIGS, edge_cases, load_rule, unpack_vecs

gpu = pytest.mark.gpu


def _alpha(a):
    return None if a == "None" else (int(a) if str(a).isdigit() else a)


# --------------------------------------------------------------------------- CPU
def test_utils_and_io_host_side(tmp_path):
    assert skm.utils.split_file_ext("x/y/file.fasta.gz") == ("file", "fasta")
    assert skm.utils.split_file_ext("file.faa") == ("file", "faa")
    assert skm.utils.check_list([1]) and skm.utils.check_list(np.zeros(2)) and not skm.utils.check_list(3)
    assert skm.io.define_output_dir(2, 8, nested=True) == os.path.join("output", "solvacc", "k-08")
    assert skm.io.define_output_dir("miqs", 3, nested="False") == "output"
    fa = tmp_path / "t.fasta"
    fa.write_text(">a|b|c desc\nMKV\nLA\n>second\n\nAC*\n")
    ids, seqs = skm.io.read_fasta(str(fa))
    assert ids == ["a|b|c", "second"] and seqs == ["MKVLA", "AC*"]
    res, off, lens = skm.io.pack_sequences(seqs)
    assert bytes(res) == b"MKVLAAC*" and off.tolist() == [0, 5, 8] and lens.tolist() == [5, 3]
    bt = tmp_path / "basis.txt"
    bt.write_text("AAA\nAAC\n")
    assert skm.io.read_kmers(str(bt)) == ["AAA", "AAC"]
    m = skm.utils.to_feature_matrix([[2, 4], [3, 9]], [2, 3])
    assert np.allclose(m, [[1, 2], [1, 3]])


def test_reduce_matches_reference_strings():
    fx = edge_cases()
    for key, case in fx["cases"].items():
        a = _alpha(key.split(":")[0])
        for (_, s), want in zip(fx["sequences"], case["reduced"]):
            assert skm.vectorize.reduce(s, a) == want
            assert skm.vectorize.reduce(s, a, mapping=skm.alphabet.FULL_ALPHABETS) == want


def test_alphabet_api_matches_oracle_tables():
    for a in (0, 1, 2, 3, 4, 5, "ptm", None, "None", "miqs"):
        assert skm.alphabet.get_alphabet_keys(a) == set(O.symbols_of(O.alphabet_name(a)))
    with pytest.raises(ValueError):
        skm.alphabet.check_valid("nope")
    with pytest.raises(ValueError):
        skm.alphabet.check_valid(9)
    with pytest.raises(KeyError):          # 6 passes check_valid, then fails in ALPHABET_ORDER (reference quirk)
        skm.alphabet.get_alphabet(6)
    assert skm.alphabet.ALPHABET2ID["miqs"] == "RED5" and skm.alphabet.ALPHABET_ORDER[2] == "solvacc"


def test_kmerbasis_argument_checks_and_pickle():
    kb = skm.vectorize.KmerBasis()
    with pytest.raises(TypeError):
        kb.set_basis(5)
    kb.set_basis(["AA", "AB", "ZZ"])
    assert kb.basis_order == {0: "AA", 1: "AB", 2: "ZZ"}
    with pytest.raises(TypeError):
        kb.transform([[1, 2]], 7)
    with pytest.raises(ValueError):
        kb.transform([[1, 2, 3]], ["AB", "QQ"])
    kv = skm.vectorize.KmerVec(alphabet=2, k=3)
    kv.set_kmer_set(["AAA", "CCC"])
    kv2 = pickle.loads(pickle.dumps(kv))
    assert kv2.k == 3 and kv2.alphabet == 2 and list(kv2.kmer_set.kmers) == ["AAA", "CCC"] and kv2.char_set == set("CAP")
    assert kv2.snekmer_version == "1.3.0" and kv2.vector is None
    assert sorted(skm.vectorize.KmerSet(0, 2)._kmerlist) == ["SS", "SV", "VS", "VV"]


# --------------------------------------------------------------------------- GPU
@gpu
def test_kmervec_reduce_vectorize_edge_cases():
    fx = edge_cases()
    seqs = [s for _, s in fx["sequences"]]
    for key, case in fx["cases"].items():
        a, k = key.split(":")
        a, k = _alpha(a), int(k)
        kv = skm.vectorize.KmerVec(alphabet=a, k=k)
        if len(kv.char_set) ** k >= 2 ** 64:
            with pytest.raises(skm.SkmError):
                kv.reduce_vectorize(seqs[0])
            continue
        got = kv.reduce_vectorize_batch(seqs)
        for g, want in zip(got, case["kmers"]):
            assert list(g) == want
            if not want:
                assert g.dtype == np.dtype("<U1") and g.shape == (0,)
        one = kv.reduce_vectorize(seqs[0])
        assert list(one) == case["kmers"][0]
        assert list(kv._kmer_gen(case["reduced"][0])) == case["kmers"][0]


@gpu
def test_kmerbasis_transform_probe():
    kb = skm.vectorize.KmerBasis()
    kb.set_basis(["AA", "AB", "ZZ"])
    out = kb.transform([[1, 2, 3], [4, 5, 6]], ["AB", "QQ", "AA"])
    assert out.tolist() == [[3, 1, 0], [6, 4, 0]]
    out = kb.transform(np.array([[1.5, 2.5, 3.5]]), ["AB", "QQ", "AA"])
    assert out.dtype == np.float64 and out.tolist() == [[3.5, 1.5, 0.0]]
    kv = skm.vectorize.KmerVec(alphabet=2, k=2)
    kv.set_kmer_set(["AA", "AB", "ZZ"])
    assert kv.harmonize(np.array([[7, 8, 9]]), ["ZZ", "AA", "AB"]).tolist() == [[8, 9, 7]]


def _golden_annotation():
    return [os.path.join(GOLDEN, "syn.ann")]


@gpu
@pytest.mark.parametrize("name", sorted(RULE_CONFIGS))
def test_rules_reproduce_reference_files(name, tmp_path):
    from snekmer_b200 import rules as R

    a, k, mf = RULE_CONFIGS[name]
    a = _alpha(a)
    d = load_rule(name)
    counts_files = []
    for nb in ("synA", "synB"):
        npz = str(tmp_path / f"{nb}.npz")
        kobj = str(tmp_path / f"{nb}.kmers")
        side = str(tmp_path / f"{nb}.skmv")
        R.vectorize_rule(os.path.join(GOLDEN, f"{nb}.fasta"), npz, kobj, a, k, min_filter=mf, out_sidecar=side)
        z = np.load(npz)
        # the binary side-car exports back to the same .npz content
        from snekmer_b200 import sidecar as SC
        SC.export_npz(side, str(tmp_path / f"{nb}.side.npz"))
        z2 = np.load(str(tmp_path / f"{nb}.side.npz"))
        for key in ("kmerlist", "ids", "seqs", "vecs", "lengths"):
            assert z2[key].dtype == z[key].dtype and np.array_equal(z2[key], z[key]), key
        assert sorted(z.files) == ["ids", "kmerlist", "lengths", "seqs", "vecs"]
        assert list(z["kmerlist"]) == list(d[f"{nb}_kmerlist"]) and z["kmerlist"].dtype == d[f"{nb}_kmerlist"].dtype
        assert list(z["ids"]) == list(d[f"{nb}_ids"])
        assert list(z["seqs"]) == list(d[f"{nb}_seqs"])
        assert z["lengths"].dtype == np.int64 and list(z["lengths"]) == list(d[f"{nb}_lengths"])
        assert z["vecs"].dtype == np.float64
        assert np.array_equal(z["vecs"].astype(np.uint8), unpack_vecs(d, f"{nb}_"))
        kv = skm.io.load_pickle(kobj)
        assert kv.k == k and list(kv.kmer_set.kmers) == list(d[f"{nb}_kmerlist"]) and list(kv.basis.basis) == list(d[f"{nb}_kmerlist"])
        # learn
        out = str(tmp_path / f"kmer-counts-{nb}.csv")
        R.learn_rule(npz, _golden_annotation(), out)
        assert open(out, "rb").read() == d[f"{nb}_counts_csv"].tobytes()
        counts_files.append(out)
    tot = str(tmp_path / "kmer-counts-total.csv")
    R.merge_rule(counts_files, tot)
    assert open(tot, "rb").read() == d["totals_csv"].tobytes()
    totA = str(tmp_path / "totalsA.csv")
    R.merge_rule(counts_files[:1], totA)
    assert open(totA, "rb").read() == d["totalsA_csv"].tobytes()
    # eval_apply (learn.smk:601-894) with the full association matrix kept
    for nb in ("synA", "synB"):
        out = str(tmp_path / f"seq-annotation-scores-{nb}.csv")
        r = R.eval_apply_rule(str(tmp_path / f"{nb}.npz"), _golden_annotation(), tot, out, save_associations=True)
        assert r.rows == list(d[f"{nb}_eval_rows"]) and r.annotations == list(d[f"{nb}_eval_cols"])
        assert np.max(np.abs(r.scores - d[f"{nb}_eval_scores"])) < 1e-12
        back = pd.read_csv(out, index_col="__index_level_0__")
        assert list(back.index) == r.rows and list(back.columns) == r.annotations
        assert np.allclose(back.values, d[f"{nb}_eval_scores"], rtol=0, atol=1e-12)
        # top-2 masking variant
        r2 = R.eval_apply_rule(str(tmp_path / f"{nb}.npz"), _golden_annotation(), tot, out, save_associations=False)
        ref = d[f"{nb}_eval_scores"]
        assert (np.isfinite(r2.scores).sum(axis=1) == min(2, ref.shape[1])).all()
        kept = np.isfinite(r2.scores)
        assert np.max(np.abs(r2.scores[kept] - ref[kept])) < 1e-12
    # apply: synB against the matrix learned on synA, with a synthetic confidence table
    conf = str(tmp_path / "global-confidence-scores.csv")
    with open(conf, "w") as f:
        f.write("Difference,confidence,weight,sum\n")
        for i in range(101):
            f.write(f"{i / 100:.2f},{min(1.0, 0.5 + i / 150):.6f},1,1\n")
    summ = str(tmp_path / "kmer-summary-synB.csv")
    full = str(tmp_path / "seq-annotation-scores-synB-apply.csv")
    r = R.apply_rule(str(tmp_path / "synB.npz"), totA, conf, summ, out_scores=full, save_associations=True)
    ref = d["apply_scores"]
    assert r.rows == list(d["apply_rows"]) and r.annotations == list(d["apply_cols"])
    assert np.max(np.abs(r.scores - ref)) < 1e-12
    i1, i2, s1, s2 = O.top2(ref)
    clear = (s1 - s2) > 1e-9
    assert np.array_equal(r.top1[clear], i1[clear])
    assert np.allclose(r.score1, s1, rtol=1e-5, atol=0)
    table = pd.read_csv(summ)
    assert list(table.columns) == ["index", "Prediction", "Score", "delta", "Confidence"]
    assert list(table["index"]) == r.rows
    assert [str(x) for x in table["Prediction"][clear]] == [str(d["apply_cols"][i]) for i in i1[clear]]
    assert np.allclose(table["Score"], s1, rtol=1e-5, atol=0)
    want_delta = np.round(s1 - s2, 2)
    near_boundary = np.abs(((s1 - s2) * 100) % 1 - 0.5) < 1e-7
    assert np.array_equal(table["delta"].values[~near_boundary], want_delta[~near_boundary])
    conf_map = R.read_confidence_csv(conf)
    want_conf = np.array([conf_map[float(x)] for x in table["delta"].values])
    assert np.allclose(table["Confidence"].values, want_conf)


@gpu
def test_vectorize_rule_with_basis_txt(tmp_path):
    from snekmer_b200 import rules as R

    d = np.load(os.path.join(GOLDEN, "basis_txt.npz"))
    bt = tmp_path / "basis.txt"
    bt.write_text("\n".join(d["basis"]) + "\n")
    npz = str(tmp_path / "synA.npz")
    R.vectorize_rule(os.path.join(GOLDEN, "synA.fasta"), npz, None, 2, 3, min_filter=5, basis_file=str(bt))
    z = np.load(npz)
    assert list(z["kmerlist"]) == list(d["kmerlist"])
    shape = tuple(d["vecs_shape"])
    want = np.unpackbits(d["vecs_bits"])[: int(np.prod(shape))].reshape(shape)
    assert np.array_equal(z["vecs"].astype(np.uint8), want)


@gpu
def test_learn_duplicate_ids_and_missing_pipes():
    from snekmer_b200 import rules as R

    ids = ["tr|A1|x", "tr|A2|y", "tr|A1|x", "tr|A3|z"]
    seqs = ["AAAC", "CCCA", "ACAC", "AAAA"]
    kmerlist = ["AA", "AC", "CA", "CC"]
    ann = {"A1": "F1", "A3": "F2"}
    r = R.learn_counts(ids, seqs, kmerlist, ann)
    # oracle on the same dict semantics
    counts = np.array([O.counts_str(s, kmerlist, None) for s in seqs])
    anns, M, nseq, totals, total = O.learn_matrix(ids, counts, ann)
    assert r.annotations == anns and np.array_equal(r.M, M) and np.array_equal(r.seq_count, nseq)
    assert np.array_equal(r.totals, totals) and r.total_seqs == total == 3
    with pytest.raises(IndexError):
        R.learn_counts(["nopipes"], ["AAAA"], kmerlist, ann)


@gpu
def test_apply_compare_check_exits(tmp_path):
    from snekmer_b200 import rules as R

    t = R.CountsTable(["Totals", "F1"], [f"{a}{b}" for a in "AC" for b in "ACDF"] + ["AA"] * 4,
                      np.array([2, 2]), np.array([5, 5]), np.ones((2, 12), dtype=np.int64))
    p = str(tmp_path / "tot.csv")
    R.write_totals_csv(p, t)
    np.savez_compressed(str(tmp_path / "q.npz"), kmerlist=np.array(["AAA"] * 12), ids=["a"], seqs=["AAAA"],
                        vecs=np.zeros((1, 12)), lengths=[4])
    with pytest.raises(SystemExit):
        R.apply_rule(str(tmp_path / "q.npz"), p, p, str(tmp_path / "o.csv"))
