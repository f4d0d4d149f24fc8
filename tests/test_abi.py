"""CPU-only: the C-ABI library builds, loads and exports every symbol declared
in include/skm_b200.h; host-only entry points work; compute fails loudly
without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "skm_b200.h")).read()
    return sorted(set(re.findall(r"SKM_API[^;(]*?\b(skm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from snekmer_b200 import _native, build

    build.build()
    names = _declared()
    assert len(names) >= 15
    handle = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    assert sorted(_native.SIGNATURES) == names
    assert _native.lib().skm_version() >= 10000


def test_lut_build_matches_oracle():
    from oracle import skm_oracle as O
    from snekmer_b200 import alphabet as A

    for a in (0, 1, 2, 3, 4, 5, "ptm", None, "None"):
        lut, syms = O.build_lut(a)
        assert A.symbols(a) == syms
        assert np.array_equal(np.frombuffer(A.lut(a), dtype=np.uint8), lut)
    # the two reference quirks
    hc = np.frombuffer(A.lut("hydrocharge"), dtype=np.uint8)
    assert hc[ord("E")] == 0xFF and A.symbols(3)[hc[ord("N")]] == "C"
    hs = np.frombuffer(A.lut("hydrostruct"), dtype=np.uint8)
    assert A.symbols(4)[hs[ord("B")]] == "B"


def test_compute_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from snekmer_b200 import engine as E

    with pytest.raises(E.SkmError):
        E.SequenceBatch.from_strings(["ACD"])


def test_plain_c_consumer(tmp_path):
    """include/skm_b200.h compiles as C99 and a C program drives the host-only entry points through the .so."""
    import shutil
    import subprocess

    from snekmer_b200 import _native, build

    build.build()
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = str(tmp_path / "host_only")
    libdir = os.path.dirname(_native.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cabi", "host_only.c"), "-o", exe, "-L", libdir, "-lskm_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "cabi ok" in r.stdout
