"""The native FASTA parser (host C++ in libskm_b200.so; no GPU needed) against the Python reader that restates
Bio.SeqIO "fasta" semantics, on the bundled fixtures and on hostile inputs, for every thread count."""
import gzip
import os

import numpy as np
import pytest

from snekmer_b200 import io as skio
from util import GOLDEN


def _want(text: bytes, tmp_path):
    p = tmp_path / "x.fasta"
    p.write_bytes(text)
    return skio.read_fasta(str(p))


def _check(text: bytes, tmp_path, threads):
    ids, seqs = _want(text, tmp_path)
    gi, res, off = skio.parse_fasta_bytes(text, threads)
    assert gi == ids
    assert off[0] == 0 and len(off) == len(seqs) + 1
    raw = res.tobytes().decode("latin-1")
    assert [raw[off[i]:off[i + 1]] for i in range(len(seqs))] == seqs


@pytest.mark.parametrize("threads", [1, 2, 3, 7, 16])
def test_fixtures(threads, tmp_path):
    for name in ("synA.fasta", "synB.fasta"):
        _check(open(os.path.join(GOLDEN, name), "rb").read(), tmp_path, threads)


@pytest.mark.parametrize("threads", [1, 4])
def test_hostile_inputs(threads, tmp_path):
    cases = [
        b"", b"\n\n", b"no record here\nACDE\n", b">only_header", b">only_header\n", b">a\nACD", b">a\nACD\n>b\n\n>c\nEF\nGH\n",
        b"junk before\n>a desc ription\nAC DE\r\nFG\t \r\n\n>b|x|y more\nKL*\n>\nMN\n> spaced title\nPQ\n",
        b">a\nAC>DE\n>b\nFG\n",                 # '>' inside a line is sequence text
        b">a\r\nACDE\r\n>b\r\nFGHI\r\n",         # CRLF
        b">a\n" + b"ACDEFGHIKL\n" * 1000 + b">b\n" + b"MNPQ\n" * 10,
    ]
    for c in cases:
        _check(c, tmp_path, threads)


def test_large_multithreaded_matches_single(tmp_path):
    rng = np.random.default_rng(0)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    parts = []
    for i in range(20000):
        L = int(rng.integers(0, 700))
        s = rng.choice(aa, size=L).tobytes()
        lines = b"\n".join(s[j:j + 60] for j in range(0, L, 60))
        parts.append(b">tr|A%07d|NAME_%d some description\n" % (i, i) + lines + b"\n")
    text = b"".join(parts)                       # ~8 MB: several ranges per thread count
    a = skio.parse_fasta_bytes(text, 1)
    for t in (2, 5, 16, 0):
        b = skio.parse_fasta_bytes(text, t)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    ids, seqs = _want(text, tmp_path)
    assert a[0] == ids and int(a[2][-1]) == sum(len(s) for s in seqs)
    # gz round trip through the file API
    p = tmp_path / "big.fasta.gz"
    with gzip.open(p, "wb", compresslevel=1) as f:
        f.write(text)
    g = skio.read_fasta_packed(str(p), 4)
    assert g[0] == ids and np.array_equal(g[1], a[1])
