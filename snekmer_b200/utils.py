"""utils: helpers kept for drop-in compatibility with ``snekmer.utils``.

Same names and behaviour as the reference helpers the hot-path callers use
(snekmer/utils.py:46-74 ``log_runtime``, :77-93 ``check_list``, :153-180
``split_file_ext``, :183-203 ``to_feature_matrix``, :206-253 ``count_n_seqs`` /
``check_n_seqs``).  Host-only; nothing here touches the device.
"""
from __future__ import annotations

import collections.abc
import datetime
import os
from typing import Any, List, Optional, Tuple

import numpy as np
import pandas as pd


def _format_timedelta(delta: datetime.timedelta) -> str:
    """'1h 41m 20.10s'-style rendering used in the rule logs."""
    ms = round(delta.microseconds / 1000)
    m, s = divmod(delta.seconds, 60)
    h, m = divmod(m, 60)
    return f"{h}h {m:02}m {s:02}.{ms}s"


def log_runtime(filename: str, start_time: datetime.datetime, step: Optional[str] = None) -> None:
    """Append start / end / total time to a rule log file."""
    end = datetime.datetime.now()
    label = "end" if step is None else step
    with open(filename, "a") as f:
        f.write(f"start time:\t{start_time}\n{label} time:\t{end}\ntotal time:\t{_format_timedelta(end - start_time)}")


def check_list(array: Any) -> bool:
    """True for sequences (incl. str, like the reference), numpy arrays and pandas Series."""
    return isinstance(array, (collections.abc.Sequence, np.ndarray, pd.Series))


def split_file_ext(filename: str) -> Tuple[str, str]:
    """('file', 'ext') for 'dir/file.ext' and for 'dir/file.ext.gz'."""
    name = os.path.basename(filename)
    stem, ext = os.path.splitext(name)
    if ext == ".gz":
        stem, ext = os.path.splitext(stem)
    return stem, ext.lstrip(".")


def to_feature_matrix(array: List, length_array=None) -> np.ndarray:
    """Rows of `array` divided by the matching entry of `length_array` (default 1)."""
    if length_array is None:
        length_array = np.ones(len(array))
    return np.asarray([np.array(row) / n for row, n in zip(array, length_array)])


def count_n_seqs(filename: str) -> int:
    """Number of FASTA records (lines starting with '>')."""
    with open(filename) as f:
        return sum(1 for line in f if line.startswith(">"))


def check_n_seqs(filename: str, k: int, show_warning: bool = True) -> bool:
    """True when the file holds at least k sequences (warns otherwise)."""
    n = count_n_seqs(filename)
    if n < k and show_warning:
        print(f"\nWARNING: {filename} contains an insufficient number of sequences for model cross-validation and will"
              f" thus be excluded from Snekmer modeling. ({k} folds specified in config; {n} sequence(s) detected.)\n")
    return n >= k
