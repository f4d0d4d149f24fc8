"""The sparse rule bodies (rules_sparse.py: learn / merge / apply on the COO + SpMM device paths with the .skmc side-car
between them) against the dense rule bodies and the CSV files written by the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, RULE_CONFIGS, load_rule

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from snekmer_b200 import engine as E
    from snekmer_b200 import io as skio
    from snekmer_b200 import rules as R
    from snekmer_b200 import rules_sparse as RS
    from snekmer_b200 import sidecar as SC


def _alpha(a):
    return None if a == "None" else (int(a) if str(a).isdigit() else a)


def _annotation():
    return [os.path.join(GOLDEN, "syn.ann")]


@pytest.mark.parametrize("name", ["hydro_k8", "miqs_k3", "solvacc_k4"])
def test_sparse_rules_match_dense_rules_and_reference_files(name, tmp_path):
    a, k, mf = RULE_CONFIGS[name]
    a = _alpha(a)
    d = load_rule(name)
    seq_annot = R.load_annotations(_annotation())
    parts, kmerlists, frames = [], [], {}
    for nb in ("synA", "synB"):
        npz = str(tmp_path / f"{nb}.npz")
        R.vectorize_rule(os.path.join(GOLDEN, f"{nb}.fasta"), npz, None, a, k, min_filter=mf)
        kl, df = skio.load_npz(npz)
        frames[nb] = df
        kmerlist = [str(x) for x in kl[0]]
        kmerlists.append(kmerlist)
        symbols = "".join(sorted(set("".join(kmerlist))))
        sc = RS.learn_counts_sparse(list(df["sequence_id"]), list(df["sequence"]), symbols, k, seq_annot)
        # dense rule on the same file
        dense = R.learn_counts(list(df["sequence_id"]), list(df["sequence"]), kmerlist, seq_annot)
        assert sc.annotations == dense.annotations and np.array_equal(sc.seq_count, dense.seq_count) and sc.total_seqs == dense.total_seqs
        codes, ok = E.encode_kmers(kmerlist, sc.symbols, k)
        assert ok.all()
        M = np.zeros((len(sc.annotations), sc.S), dtype=np.int64)
        kk = sc.keys.cpu().numpy()
        M[kk // sc.S, kk % sc.S] = sc.vals.cpu().numpy()
        assert np.array_equal(M[:, codes.astype(np.int64)], dense.M)
        assert np.array_equal(sc.totals.cpu().numpy()[codes.astype(np.int64)], dense.totals)
        # side-car in the reference's column order -> the reference's own CSV, byte for byte
        side = str(tmp_path / f"{nb}.skmc")
        RS.write_counts_sidecar(side, sc, kmers=kmerlist)
        out = tmp_path / f"kmer-counts-{nb}.csv"
        SC.export_counts_csv(side, str(out))
        assert out.read_bytes() == d[f"{nb}_counts_csv"].tobytes()
        # side-car in code order -> back to the device, unchanged
        side2 = str(tmp_path / f"{nb}.codes.skmc")
        RS.write_counts_sidecar(side2, sc)
        back = RS.read_counts_sidecar(side2, symbols=sc.symbols)
        assert back.annotations == sc.annotations and torch.equal(back.keys, sc.keys) and torch.equal(back.vals, sc.vals)
        assert torch.equal(back.totals, sc.totals) and back.total_seqs == sc.total_seqs and np.array_equal(back.seq_count, sc.seq_count)
        parts.append(sc)
    # merge == the dense merge of the two reference CSVs
    merged = RS.merge_counts_sparse(parts)
    csvs = []
    for nb in ("synA", "synB"):
        p = tmp_path / f"ref-{nb}.csv"
        p.write_bytes(d[f"{nb}_counts_csv"].tobytes())
        csvs.append(str(p))
    ref = R.merge_tables([R.read_counts_csv(c) for c in csvs])
    assert ["Totals"] + merged.annotations == ref.rows
    assert np.array_equal(np.concatenate([[merged.total_seqs], merged.seq_count]), ref.seq_count)
    codes, ok = E.encode_kmers(ref.kmers, merged.symbols, k)
    assert ok.all()
    Mm = np.zeros((len(merged.annotations), merged.S), dtype=np.int64)
    kk = merged.keys.cpu().numpy()
    Mm[kk // merged.S, kk % merged.S] = merged.vals.cpu().numpy()
    assert np.array_equal(Mm[:, codes.astype(np.int64)], ref.M[1:])
    assert np.array_equal(merged.totals.cpu().numpy()[codes.astype(np.int64)], ref.M[0])
    # apply: synB against the matrix learned on synA — the reference's own cosine matrix decides
    df = frames["synB"]
    conf = str(tmp_path / "global-confidence-scores.csv")
    with open(conf, "w") as f:
        f.write("Difference,confidence,weight,sum\n")
        for i in range(101):
            f.write(f"{i / 100:.2f},{min(1.0, 0.5 + i / 150):.6f},1,1\n")
    summ = str(tmp_path / "kmer-summary-synB.csv")
    r = RS.apply_counts_sparse(list(df["sequence_id"]), list(df["sequence"]), parts[0], conf, summ)
    S = d["apply_scores"]
    assert r.rows == list(d["apply_rows"]) and r.annotations == list(d["apply_cols"])
    order = np.argsort(-S, axis=1, kind="stable")
    s1 = S[np.arange(len(S)), order[:, 0]]
    s2 = S[np.arange(len(S)), order[:, 1]]
    assert np.max(np.abs(r.score1 - s1)) < 1e-12 and np.max(np.abs(r.score2 - s2)) < 1e-12
    clear = (s1 - s2) > 1e-9
    assert np.array_equal(r.top1[clear], order[:, 0][clear])
    import pandas as pd

    table = pd.read_csv(summ)
    assert list(table.columns) == ["index", "Prediction", "Score", "delta", "Confidence"] and list(table["index"]) == r.rows
    assert [str(x) for x in table["Prediction"][clear]] == [str(d["apply_cols"][i]) for i in order[:, 0][clear]]
    dense_r = R.cosine_top2(list(df["sequence"]), kmerlists[1], R.read_counts_csv(csvs[0]))
    assert np.array_equal(dense_r.top1.cpu().numpy(), r.top1) and np.max(np.abs(dense_r.score1.cpu().numpy() - r.score1)) < 1e-12
