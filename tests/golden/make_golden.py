#!/usr/bin/env python
"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Inputs are synthetic and seeded (written next to the outputs so the tests do
not need the reference); outputs are what the reference's own modules and rule
bodies (driven through tests/refharness.py) produce for them:

  edge_kmers.json.gz      KmerVec.reduce_vectorize / reduce on hand-written edge
                          cases, every built-in alphabet x k in {1,2,3,5,8,14}
  syn{A,B}.fasta, syn.ann inputs for the rule-level fixtures
  rule_<cfg>.npz          per config: kmerlist/ids/seqs/lengths/vecs(bit-packed)
                          of the vectorize rule for both files, the learn
                          kmer-counts CSVs, the merged totals CSV, the
                          eval_apply and apply cosine matrices (float64)
  basis_txt.npz           the basis.txt branch of the vectorize rule
  MANIFEST.json           library versions the reference ran under
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refharness as rh  # noqa: E402

AA = "ACDEFGHIKLMNPQRSTVWY"
BG = np.array([.122, .009, .060, .057, .034, .084, .021, .047, .025, .105,
               .024, .022, .053, .034, .074, .047, .050, .071, .014, .022])
BG = BG / BG.sum()

EDGE = [
    ("tr|E000|plain", "MKVLAAGIVGLLLAQPSWA"),
    ("tr|E001|trailing_star", "MKVLAAGIVGLLLAQPSWA*"),
    ("tr|E002|three_stars", "MKTAYIAKQRQISFVKSHFSRQ***"),
    ("tr|E003|internal_star", "MKTAYIAK*QRQISFVKSHFSRQ"),
    ("tr|E004|lowercase", "MKTAYIAKqrqisfVKSHFSRQLEERLGLIEVQ"),
    ("tr|E005|ambiguity", "MKXTAYBIAKZQRUQISOFVKJSHFSRQ"),
    ("tr|E006|short", "MK"),
    ("tr|E007|single", "A"),
    ("tr|E008|empty", ""),
    ("tr|E009|only_star", "*"),
    ("tr|E010|all_invalid", "XXXXXXXXXXXXXXXXXXXX"),
    ("tr|E011|homopolymer", "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAA"),
    ("tr|E012|repeat", "ACDACDACDACDACDACDACDACDACD"),
    ("tr|E013|glu_asn", "EEENNNEEENNNKDRKDREEE"),
    ("tr|E014|rawB", "LLHHBBPGPGBBLLHHBB"),
    ("tr|E015|ptm", "MK-T_A!Y^I#A$K@Q.R%Q&ISFVK"),
    ("tr|E016|exact14", "MKTAYIAKQRQISF"),
    ("tr|E017|x_every_5", "MKTAXYIAKXQRQIXSFVKXSHFSXRQLEXERLGX"),
    ("tr|E018|digits_space", "MKT1AY IAK\tQRQ"),
    ("tr|E019|star_then_more", "MKTAY**IAKQRQISFVKSHF*"),
]


def synth_fasta(path, n, seed, tag):
    rng = np.random.default_rng(seed)
    recs = []
    fams = []
    # a few "families": members are mutated copies so that cosine top-1 is meaningful
    for f in range(6):
        L = int(rng.integers(60, 260))
        fams.append(rng.choice(list(AA), size=L, p=BG))
    for i in range(n):
        if rng.random() < 0.6:
            base = fams[int(rng.integers(0, len(fams)))].copy()
            mut = rng.random(len(base)) < 0.12
            base[mut] = rng.choice(list(AA), size=int(mut.sum()), p=BG)
            s = "".join(base)
            if rng.random() < 0.3:
                cut = int(rng.integers(0, len(s) // 3))
                s = s[cut:]
        else:
            L = int(np.clip(round(rng.lognormal(5.0, 0.6)), 12, 400))
            s = "".join(rng.choice(list(AA), size=L, p=BG))
        r = rng.random()
        if r < 0.08:
            p = int(rng.integers(0, len(s)))
            s = s[:p] + "X" + s[p:]
        elif r < 0.12:
            s = s + "*"
        elif r < 0.14:
            s = s[:5]
        recs.append((f"tr|{tag}{i:04d}|{tag}{i:04d}_SYNTH", s))
    with open(path, "w") as f:
        for rid, s in recs:
            f.write(f">{rid} synthetic protein OS=none OX=0\n")
            for j in range(0, len(s), 60):
                f.write(s[j:j + 60] + "\n")
    return recs


def write_ann(path, recs, seed, nfam):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        f.write("id\tTIGRFAMs\n")
        for rid, _ in recs:
            acc = rid.split("|")[1]
            if rng.random() < 0.7:
                f.write(f"{acc}\tTIGR{int(rng.integers(0, nfam)):05d}\n")


def edge_fixture(skm):
    out = {}
    alphabets = [0, 1, 2, 3, 4, 5, "ptm", "None"]
    for a in alphabets:
        for k in (1, 2, 3, 5, 8, 14):
            kv = skm.vectorize.KmerVec(alphabet=a if a != "None" else None, k=k)
            key = f"{a}:{k}"
            out[key] = {
                "kmers": [list(map(str, kv.reduce_vectorize(s))) for _, s in EDGE],
                "reduced": [skm.vectorize.reduce(s, alphabet=a if a != "None" else None,
                                                 mapping=skm.alphabet.FULL_ALPHABETS)
                            for _, s in EDGE],
                "char_set": sorted(kv.char_set),
            }
    return {"sequences": EDGE, "cases": out}


RULE_CONFIGS = {
    # name: (alphabet, k, min_filter)
    "solvacc_k4": (2, 4, None),
    "miqs_k3": (5, 3, None),
    "hydro_k8": (0, 8, None),
    "none_k2": ("None", 2, None),
    "standard_k5_mf1": ("standard", 5, 1),
    "hydrocharge_k6": (3, 6, None),
}


def rule_fixture(name, alphabet, k, min_filter, work):
    cfg = {"alphabet": alphabet if alphabet != "None" else None, "k": k,
           "learnapp": {"save_apply_associations": True}}
    if min_filter is not None:
        cfg["min_filter"] = min_filter
    wd = os.path.join(work, name)
    os.makedirs(wd)
    ann = os.path.join(HERE, "syn.ann")
    out = {}
    counts_csv = []
    for nb in ("synA", "synB"):
        npz = rh.run_vectorize(wd, os.path.join(HERE, f"{nb}.fasta"), nb, cfg)
        d = np.load(npz)
        out[f"{nb}_kmerlist"] = d["kmerlist"]
        out[f"{nb}_ids"] = d["ids"]
        out[f"{nb}_seqs"] = d["seqs"]
        out[f"{nb}_lengths"] = d["lengths"]
        v = d["vecs"]
        assert set(np.unique(v)) <= {0.0, 1.0}
        out[f"{nb}_vecs_shape"] = np.array(v.shape)
        out[f"{nb}_vecs_bits"] = np.packbits(v.astype(np.uint8), axis=None)
        c = rh.run_learn(wd, nb, [ann], cfg)
        counts_csv.append(c)
        out[f"{nb}_counts_csv"] = np.frombuffer(open(c, "rb").read(), dtype=np.uint8)
    tot = rh.run_merge(wd, counts_csv, cfg)
    out["totals_csv"] = np.frombuffer(open(tot, "rb").read(), dtype=np.uint8)
    for nb in ("synA", "synB"):
        s = rh.run_eval_apply_scores(wd, nb, [ann], tot, cfg)
        out[f"{nb}_eval_scores"] = s.values.astype(np.float64)
        out[f"{nb}_eval_rows"] = np.array(list(s.index), dtype=str)
        out[f"{nb}_eval_cols"] = np.array(list(s.columns), dtype=str)
    # apply: synB as queries against a matrix learned from synA only
    totA = rh.run_merge(wd, counts_csv[:1], cfg)
    out["totalsA_csv"] = np.frombuffer(open(totA, "rb").read(), dtype=np.uint8)
    s = rh.run_apply_scores(wd, "synB", totA, cfg)
    out["apply_scores"] = s.values.astype(np.float64)
    out["apply_rows"] = np.array(list(s.index), dtype=str)
    out["apply_cols"] = np.array(list(s.columns), dtype=str)
    np.savez_compressed(os.path.join(HERE, f"rule_{name}.npz"), **out)


def basis_txt_fixture(work):
    """kmerize.smk:72-78 branch: a supplied (sorted) basis.txt."""
    cfg = {"alphabet": 2, "k": 3}
    wd = os.path.join(work, "basis_txt")
    os.makedirs(wd)
    skm = rh.load_reference()
    kv = skm.vectorize.KmerVec(alphabet=2, k=3)
    basis = sorted("".join(p) for p in __import__("itertools").product(sorted(kv.char_set), repeat=3))
    basis = basis[::2] + ["ZZZ"]          # a subset + one k-mer that never occurs
    bt = os.path.join(wd, "basis.txt")
    with open(bt, "w") as f:
        f.write("\n".join(basis) + "\n")
    npz = rh.run_vectorize(wd, os.path.join(HERE, "synA.fasta"), "synA", cfg, basis_txt=bt)
    d = np.load(npz)
    np.savez_compressed(os.path.join(HERE, "basis_txt.npz"), basis=np.array(basis),
                        kmerlist=d["kmerlist"], vecs_shape=np.array(d["vecs"].shape),
                        vecs_bits=np.packbits(d["vecs"].astype(np.uint8), axis=None))


def main():
    import pandas
    import pyarrow
    import sklearn
    skm = rh.load_reference()
    recsA = synth_fasta(os.path.join(HERE, "synA.fasta"), 90, 11, "A")
    recsB = synth_fasta(os.path.join(HERE, "synB.fasta"), 70, 12, "B")
    # synB re-uses a few synA accessions' families through the shared .ann
    write_ann(os.path.join(HERE, "syn.ann"), recsA + recsB, 13, 7)
    with gzip.open(os.path.join(HERE, "edge_kmers.json.gz"), "wt") as f:
        json.dump(edge_fixture(skm), f)
    work = tempfile.mkdtemp(prefix="golden_")
    try:
        for name, (a, k, mf) in RULE_CONFIGS.items():
            rule_fixture(name, a, k, mf, work)
            print("done", name)
        basis_txt_fixture(work)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump({"reference": "PNNL-CompBio/Snekmer", "version": skm._version.__version__,
                   "numpy": np.__version__, "pandas": pandas.__version__,
                   "sklearn": sklearn.__version__, "pyarrow": pyarrow.__version__,
                   "python": sys.version.split()[0],
                   "generator": "tests/golden/make_golden.py"}, f, indent=1)


if __name__ == "__main__":
    main()
