"""Signal input for evaluation: parse -> normalise -> sliding windows.

Mirrors the eval side of chiron/chiron_input.py (read_signal :527-539, read_signal_fast5 :541-555,
read_data_for_eval :253-292, padding :681-692, DataSet.next_batch(shuffle=False) :194-250) with numpy arrays instead of
Python lists.  Normalisation is a per-model property stored in the weight blob (SURVEY.md finding 4): DNA_default needs
the unique-value median/MAD that read_signal_fast5's MEDIAN branch computes.

The three per-read hot spots -- the token loop of read_signal, the np.unique sort behind the normalisation and the
per-window slicing -- run in the native library's host-side helpers (cb_host_parse_signal / cb_host_normalize /
cb_host_windows, csrc/cb_host_signal.cu): single C passes that release the GIL, so the reader threads of
chiron_eval.evaluation() work in parallel.  They are bit-identical to the numpy formulation the oracle keeps
(tests/test_host_signal.py)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .model import NORM_FULL_MAD, NORM_NONE, NORM_UNIQUE_MAD

MEAN, MEDIAN = "mean", "median"            # chiron/chiron_input.py:33-34


def parse_signal_text(data: bytes) -> np.ndarray:
    """Whitespace-separated numbers -> float32 samples (the token loop of chiron_input.py:527-532)."""
    lib = _lib.load()
    cap = len(data) // 2 + 1                # every sample takes at least one character and one separator
    out = np.empty(cap, dtype=np.float32)
    n = lib.cb_host_parse_signal(data, len(data), out.ctypes.data_as(ctypes.c_void_p), cap)
    if n < 0:
        _lib.check(int(n), "cb_host_parse_signal")
    return out[:n].copy()


def read_signal(file_path: str) -> np.ndarray:
    """Whitespace-separated samples of a ``.signal`` file (chiron_input.py:527-532), as float32."""
    with open(file_path, "rb") as f:
        return parse_signal_text(f.read())


def read_signal_fast5(fast5_path: str) -> np.ndarray:
    """First read's ``Raw/Reads/*/Signal`` (chiron_input.py:546-547) through the dependency-free HDF5 reader."""
    from .fast5 import read_raw_signal
    return read_raw_signal(fast5_path).astype(np.float32)


def normalize_signal(signal: np.ndarray, mode: int) -> np.ndarray:
    """mode NORM_UNIQUE_MAD: (s - median(unique(s))) / mad(unique(s))   (chiron_input.py:548,553-554)
    mode NORM_FULL_MAD:   (s - median(s)) / mad(s)                       (chiron_input.py:537-538)
    with mad = statsmodels.robust.mad = median(|x - median|) / 0.6745; float64 statistics, float32 result."""
    if mode not in (NORM_NONE, NORM_UNIQUE_MAD, NORM_FULL_MAD):
        raise ValueError("unknown signal normalisation %r" % (mode,))
    s = np.ascontiguousarray(signal, dtype=np.float32)
    out = np.empty_like(s)
    _lib.check(_lib.load().cb_host_normalize(s.ctypes.data_as(ctypes.c_void_p), s.size, int(mode),
                                             out.ctypes.data_as(ctypes.c_void_p)), "cb_host_normalize")
    return out


class DataSet:
    """The windows of one read.  ``next_batch`` keeps the reference's sequential, no-shuffle contract."""

    def __init__(self, event: np.ndarray, event_length: np.ndarray, samples: int = 0):
        self.event = event                  # [n, seg_length] float32, zero padded
        self.event_length = event_length    # [n] int32
        self.samples = int(samples)         # raw samples the windows were cut from (perf report)
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
    sig = np.ascontiguousarray(f_signal, dtype=np.float32)
    lib = _lib.load()
    n_win = -(-sig.size // int(step)) if step > 0 else 0
    event = np.empty((n_win, seg_length), dtype=np.float32)
    lens = np.empty(n_win, dtype=np.int32)
    n = lib.cb_host_windows(sig.ctypes.data_as(ctypes.c_void_p), sig.size, int(step), int(seg_length),
                            event.ctypes.data_as(ctypes.c_void_p), lens.ctypes.data_as(ctypes.c_void_p), n_win)
    if n < 0:
        _lib.check(int(n), "cb_host_windows")
    return DataSet(event, lens, samples=sig.size)
