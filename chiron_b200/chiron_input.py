"""Signal input for evaluation: parse -> normalise -> sliding windows.

Mirrors the eval side of chiron/chiron_input.py (read_signal :527-539, read_signal_fast5 :541-555,
read_data_for_eval :253-292, padding :681-692, DataSet.next_batch(shuffle=False) :194-250) with numpy arrays instead of
Python lists.  Normalisation is a per-model property stored in the weight blob (SURVEY.md finding 4): DNA_default needs
the unique-value median/MAD that read_signal_fast5's MEDIAN branch computes."""
from __future__ import annotations

import numpy as np

from .model import NORM_FULL_MAD, NORM_NONE, NORM_UNIQUE_MAD

MEAN, MEDIAN = "mean", "median"            # chiron/chiron_input.py:33-34
_MAD_C = 0.6744897501960817                # statsmodels.robust.mad: median(|x - median|) / Phi^-1(3/4)


def read_signal(file_path: str) -> np.ndarray:
    """Whitespace-separated samples of a ``.signal`` file (chiron_input.py:527-532), as float32."""
    with open(file_path, "r") as f:
        return np.asarray(f.read().split(), dtype=np.float32)


def read_signal_fast5(fast5_path: str) -> np.ndarray:
    """First read's ``Raw/Reads/*/Signal`` (chiron_input.py:546-547) through the dependency-free HDF5 reader."""
    from .fast5 import read_raw_signal
    return read_raw_signal(fast5_path).astype(np.float32)


def normalize_signal(signal: np.ndarray, mode: int) -> np.ndarray:
    """mode NORM_UNIQUE_MAD: (s - median(unique(s))) / mad(unique(s))   (chiron_input.py:548,553-554)
    mode NORM_FULL_MAD:   (s - median(s)) / mad(s)                       (chiron_input.py:537-538)."""
    s = np.asarray(signal, dtype=np.float64)
    if mode == NORM_NONE or s.size == 0:
        return s.astype(np.float32)
    if mode == NORM_UNIQUE_MAD:
        ref = np.unique(s)
    elif mode == NORM_FULL_MAD:
        ref = s
    else:
        raise ValueError("unknown signal normalisation %r" % (mode,))
    med = np.median(ref)
    mad = np.median(np.abs(ref - med)) / _MAD_C
    return ((s - med) / mad).astype(np.float32)


class DataSet:
    """The windows of one read.  ``next_batch`` keeps the reference's sequential, no-shuffle contract."""

    def __init__(self, event: np.ndarray, event_length: np.ndarray):
        self.event = event                  # [n, seg_length] float32, zero padded
        self.event_length = event_length    # [n] int32
        self._index = 0
        self.epochs_completed = 0

    @property
    def reads_n(self) -> int:
        return int(self.event.shape[0])

    def next_batch(self, batch_size: int, shuffle: bool = False):
        if shuffle:
            raise ValueError("evaluation data is never shuffled (chiron_eval.py:322-323)")
        start = self._index
        end = min(start + batch_size, self.reads_n)
        if start + batch_size >= self.reads_n:          # chiron_input.py:215-226: the rest of the read, epoch done
            self.epochs_completed += 1
            self._index = 0
        else:
            self._index = end
        return self.event[start:end], self.event_length[start:end], []


def read_data_for_eval(file_path: str, start_index: int = 0, step: int = 20, seg_length: int = 200,
                       reverse_fast5: bool = False, sig_norm: int = NORM_UNIQUE_MAD) -> DataSet:
    """chiron_input.py:253-292.  Windows start at 0, step, 2*step ... < n; the tail windows are zero padded and carry
    their true length."""
    if file_path.endswith(".signal"):
        f_signal = read_signal(file_path)
    elif file_path.endswith(".fast5"):
        f_signal = read_signal_fast5(file_path)
        if reverse_fast5:
            f_signal = f_signal[::-1]
    else:
        raise TypeError("Input file should be a signal file or fast5 file, but a %s file is given." % file_path)
    f_signal = normalize_signal(f_signal, sig_norm)[start_index:]
    return windows_from_signal(f_signal, step, seg_length)


def windows_from_signal(f_signal: np.ndarray, step: int, seg_length: int) -> DataSet:
    n = int(f_signal.shape[0])
    starts = np.arange(0, n, step, dtype=np.int64)
    event = np.zeros((len(starts), seg_length), dtype=np.float32)
    lens = np.minimum(seg_length, n - starts).astype(np.int32)
    if len(starts):
        idx = starts[:, None] + np.arange(seg_length, dtype=np.int64)[None, :]
        valid = idx < n
        event[valid] = f_signal[idx[valid]]
    return DataSet(event, lens)
