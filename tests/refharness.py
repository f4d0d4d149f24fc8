"""Test-only driver for the *unmodified* reference (PNNL-CompBio/Snekmer 1.3.0).

Nothing in here is product code and nothing from the reference is copied: the
reference's own modules are imported from ``/root/reference`` (behind a stub
package, because ``snekmer/__init__.py`` pulls matplotlib/hdbscan which are not
installed) and the ``run:`` bodies of its Snakemake rules are read from the
``.smk`` files *by line range at run time* and exec'd.

It is used for two things only:

* ``tests/golden/make_golden.py`` – generate the committed golden vectors;
* ``tests/test_reference_live.py`` – extra parity tests that run only where
  ``/root/reference`` exists (this container; never the GPU box).

Line ranges (reference file:line):
  kmerize.smk:67-142   vectorize rule body
  learn.smk:247-422    class Library
  learn.smk:443-594    class Merge
  learn.smk:628-887    class KmerCompare (eval_apply)
  apply.smk:147-353    class KmerCompare (apply)
  learn.smk:923-1348   class Evaluator (evaluate rule)
"""
from __future__ import annotations

import importlib
import io
import os
import sys
import types
from contextlib import contextmanager, redirect_stdout
from types import SimpleNamespace

REF_ROOT = os.environ.get("SNEKMER_REFERENCE", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "snekmer")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "vectorize.py"))


_skm = None


def load_reference():
    """Import reference sub-modules behind a stub ``snekmer`` package."""
    global _skm
    if _skm is not None:
        return _skm
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    # The stub stays registered: the vectorize rule pickles its KmerVec
    # (kmerize.smk:141-142), which needs ``snekmer.vectorize`` importable.
    # The product package is ``snekmer_b200`` so there is no name clash.
    stub = types.ModuleType("snekmer")
    stub.__path__ = [REF_PKG]
    sys.modules["snekmer"] = stub
    for name in ("_version", "alphabet", "utils", "vectorize", "io"):
        setattr(stub, name, importlib.import_module(f"snekmer.{name}"))
    _skm = stub
    return stub


# ---------------------------------------------------------------------------
# FASTA reading with Bio.SeqIO semantics as used at kmerize.smk:90-129
# (record.id = header up to first whitespace, record.seq = joined lines)
# ---------------------------------------------------------------------------
class _Rec:
    __slots__ = ("id", "seq")

    def __init__(self, rid, seq):
        self.id = rid
        self.seq = seq


class _SeqIO:
    @staticmethod
    def parse(path, fmt="fasta"):
        rid, chunks = None, []
        with open(path) as f:
            for line in f:
                line = line.rstrip("\n").rstrip("\r")
                if line.startswith(">"):
                    if rid is not None:
                        yield _Rec(rid, "".join(chunks))
                    parts = line[1:].split(None, 1)
                    rid = parts[0] if parts else ""
                    chunks = []
                elif rid is not None:
                    chunks.append(line.strip())
        if rid is not None:
            yield _Rec(rid, "".join(chunks))


def _rule_source(smk: str, first: int, last: int, indent: int = 8) -> str:
    with open(os.path.join(REF_PKG, "rules", smk)) as f:
        lines = f.readlines()[first - 1:last]
    out = []
    pad = " " * indent
    for ln in lines:
        out.append(ln[indent:] if ln.startswith(pad) else ln.lstrip(" ") if ln.strip() == "" else ln)
    return "".join(out)


@contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def _base_globals(config):
    import itertools
    import pickle
    import re
    from datetime import datetime
    from os.path import basename, exists, join

    import numpy as np
    import pandas as pd
    import pyarrow as pa
    import pyarrow.csv as pacsv
    import sklearn
    import sklearn.metrics

    skm = load_reference()
    return dict(
        skm=skm, np=np, pd=pd, pa=pa, csv=pacsv, sklearn=sklearn, re=re, sys=sys,
        itertools=itertools, pickle=pickle, datetime=datetime, basename=basename,
        exists=exists, join=join, SeqIO=_SeqIO, config=config,
    )


def run_vectorize(workdir, fasta, nb, config, basis_txt=None):
    """kmerize.smk:67-142 on one FASTA → output/vector/{nb}.npz (+ .kmers)."""
    os.makedirs(os.path.join(workdir, "output", "vector"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "output", "kmerize"), exist_ok=True)
    g = _base_globals(config)
    inp = SimpleNamespace(fasta=os.path.abspath(fasta))
    if basis_txt is not None:
        inp.kmerbasis = os.path.abspath(basis_txt)
    g["input"] = inp
    g["output"] = SimpleNamespace(
        data=os.path.join("output", "vector", f"{nb}.npz"),
        kmerobj=os.path.join("output", "kmerize", f"{nb}.kmers"),
    )
    src = _rule_source("kmerize.smk", 67, 142)
    with _cwd(workdir):
        exec(compile(src, "kmerize.smk[67:142]", "exec"), g)
    return os.path.join(workdir, "output", "vector", f"{nb}.npz")


def run_learn(workdir, nb, annotation_files, config):
    """learn.smk:247-422 (Library) → output/learn/kmer-counts-{nb}.csv."""
    os.makedirs(os.path.join(workdir, "output", "learn"), exist_ok=True)
    g = _base_globals(config)
    g["log"] = [os.path.abspath(os.path.join(workdir, "learn.log"))]
    g["start_time"] = g["datetime"].now()
    src = _rule_source("learn.smk", 247, 422)
    annotation_files = [os.path.abspath(a) for a in annotation_files]
    with _cwd(workdir), redirect_stdout(io.StringIO()):
        exec(compile(src, "learn.smk[247:422]", "exec"), g)
        lib = g["Library"]()
        # Library methods resolve skm/pd/... through the exec globals
        lib.execute_all(annotation_files,
                        f"output/vector/{nb}.npz")
    return os.path.join(workdir, "output", "learn", f"kmer-counts-{nb}.csv")


def run_merge(workdir, counts_files, config, base_counts=""):
    """learn.smk:443-594 (Merge) → output/learn/kmer-counts-total.csv."""
    g = _base_globals(config)
    src = _rule_source("learn.smk", 443, 594)
    counts_files = [os.path.abspath(c) for c in counts_files]
    if base_counts:
        base_counts = os.path.abspath(base_counts)
    out = os.path.join("output", "learn", "kmer-counts-total.csv")
    with _cwd(workdir), redirect_stdout(io.StringIO()):
        exec(compile(src, "learn.smk[443:594]", "exec"), g)
        g["Merge"](list(counts_files), base_counts, out).execute_all()
    return os.path.join(workdir, out)


def run_eval_apply_scores(workdir, nb, annotation_files, totals_csv, config):
    """learn.smk:628-829: cosine matrix of the eval_apply rule, *before* the
    top-2 mask (learn.smk:848 raises under pandas 3 copy-on-write).  Returns the
    DataFrame (rows tagged known/unknown, columns = annotations)."""
    g = _base_globals(config)
    src = _rule_source("learn.smk", 628, 887)
    annotation_files = [os.path.abspath(a) for a in annotation_files]
    totals_csv = os.path.abspath(totals_csv)
    with _cwd(workdir), redirect_stdout(io.StringIO()):
        exec(compile(src, "learn.smk[628:887]", "exec"), g)
        kc = g["KmerCompare"](totals_csv, annotation_files,
                              f"output/vector/{nb}.npz", "unused.csv")
        kc.generate_inputs()
        kc.generate_kmer_counts()
        kc.add_known_unknown_tag()
        kmer_counts = kc.construct_kmer_counts_dataframe()
        kmer_counts = kc.match_kmer_counts_format(kmer_counts)
        return kc.calculate_cosine_similarity(kmer_counts)


def run_apply_scores(workdir, nb, totals_csv, config):
    """apply.smk:147-289: the apply rule up to and including the cosine matrix
    (apply.smk:317-319 raises under pandas 3, so the summary table is restated
    in oracle/ and pinned on this matrix).  Returns DataFrame Q x A."""
    g = _base_globals(config)
    src = _rule_source("apply.smk", 147, 353)
    totals_csv = os.path.abspath(totals_csv)
    with _cwd(workdir), redirect_stdout(io.StringIO()):
        exec(compile(src, "apply.smk[147:353]", "exec"), g)
        kc = g["KmerCompare"](totals_csv, f"output/vector/{nb}.npz",
                              "unused-conf.csv", "unused-seqann.csv", "unused-summary.csv")
        kc.load_data()
        kc.generate_kmer_counts()
        kc.construct_kmer_counts_dataframe()
        kc.match_kmer_counts_format()
        kc.cosine_similarity()
        return kc.kmer_count_totals


def run_evaluate(workdir, score_csvs, out_conf, out_glob, base_confidence=(), modifier=1.0):
    """learn.smk:923-1348 (class Evaluator): seq-annotation-scores CSVs -> confidence-matrix.csv and
    global-confidence-scores.csv, optionally merged with ONE prior global-confidence file."""
    g = _base_globals({})
    g["params"] = SimpleNamespace(modifer=modifier)
    src = _rule_source("learn.smk", 923, 1348)
    with _cwd(workdir), redirect_stdout(io.StringIO()):
        exec(compile(src, "learn.smk[923:1348]", "exec"), g)
        ev = g["Evaluator"]([os.path.abspath(p) for p in score_csvs], out_conf, out_glob,
                            [os.path.abspath(p) for p in base_confidence])
        ev.execute_all()
    return os.path.join(workdir, out_conf), os.path.join(workdir, out_glob)
