"""Shared helpers for the tests (FASTA / fixture loading)."""
import gzip
import io
import json
import os

import numpy as np
import pandas as pd

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_fasta(path):
    """(ids, seqs) with Bio.SeqIO semantics (id = header up to whitespace)."""
    ids, seqs, cur = [], [], None
    with open(path) as f:
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


def read_ann(path):
    df = pd.read_table(path)
    return dict(zip(df["id"].tolist(), df["TIGRFAMs"].tolist()))


def load_rule(name):
    return np.load(os.path.join(GOLDEN, f"rule_{name}.npz"))


def csv_frame(raw: np.ndarray) -> pd.DataFrame:
    """kmer-counts CSV bytes → DataFrame (blank = 0), index = row labels."""
    df = pd.read_csv(io.BytesIO(raw.tobytes()), index_col="__index_level_0__", header=0)
    return df.fillna(0)


def unpack_vecs(d, prefix):
    shape = tuple(d[f"{prefix}vecs_shape"])
    n = int(np.prod(shape))
    return np.unpackbits(d[f"{prefix}vecs_bits"])[:n].reshape(shape)


def edge_cases():
    with gzip.open(os.path.join(GOLDEN, "edge_kmers.json.gz"), "rt") as f:
        return json.load(f)


RULE_CONFIGS = {
    "solvacc_k4": (2, 4, 0),
    "miqs_k3": (5, 3, 0),
    "hydro_k8": (0, 8, 0),
    "none_k2": ("None", 2, 0),
    "standard_k5_mf1": ("standard", 5, 1),
    "hydrocharge_k6": (3, 6, 0),
}
