"""snekmer_b200 — B200-native hot path of Snekmer behind Snekmer's own Python API.

``alphabet``, ``vectorize``, ``io`` and ``utils`` mirror the reference modules of the
same names; ``engine`` / ``pipeline`` / ``rules`` / ``dist`` are the batch layers on top
of the C-ABI CUDA library (``include/skm_b200.h``).  Importing the package does not
touch CUDA; the first compute call does, and raises ``SkmError`` without a GPU.
"""
from ._version import __b200_version__, __version__  # noqa: F401
from . import alphabet, io, utils, vectorize  # noqa: F401
from ._native import SkmError  # noqa: F401
