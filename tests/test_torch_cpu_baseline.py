"""The torch-CPU restatement that bench.py times as the CPU baseline (oracle/torch_cpu.py) against the numpy oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import chiron_oracle as O


def _windows(cfg, path, L, jump, n):
    from chiron_b200 import fast5
    if path.endswith(".fast5"):
        sig = fast5.read_raw_signal(path)[::-1].astype(np.float32)
    else:
        sig = O.read_signal_text(path)
    x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), L, jump)
    return x[-n:].copy(), lens[-n:].copy()


def test_dna_default_matches_the_oracle(dna_model):
    """Full and ragged windows of read1 (incl. a 1-sample and a 0-sample row): logits within fp32 noise of the numpy oracle,
    greedy bases identical; the per-stage timer returns the same logits."""
    from oracle.torch_cpu import TorchCpuModel
    cfg, t, _ = dna_model
    x, lens = _windows(cfg, os.path.join(GOLDEN, "DNA", "raw", "read1.signal"), 300, 290, 14)   # last window is ragged
    lens[2], lens[5] = 1, 0
    x[2, 1:] = 0
    x[5] = 0
    m = TorchCpuModel(cfg, t)
    ref = O.inference(x, lens, cfg, t)
    got = m.inference(x, lens)
    assert np.abs(got - ref).max() < 2e-3
    assert O.ctc_greedy(got, lens) == O.ctc_greedy(ref, lens)
    full = lens == 300
    times, lg, paths = m.timed_pass(x[full], lens[full], lambda a, b: O.ctc_decode_c(a, b, 0))
    assert np.abs(lg - ref[full]).max() < 2e-3 and paths == O.ctc_greedy(ref[full], lens[full])
    assert set(times) == {"conv", "lstm", "head", "decode", "total"} and times["total"] > 0


def test_rna_default_matches_the_oracle(rna_model):
    """The stride-5 / width-13 first block and the MultiRNNCell-per-direction layout (reverse_sequence on ragged rows)."""
    from oracle.torch_cpu import TorchCpuModel
    cfg, t, _ = rna_model
    x, lens = _windows(cfg, os.path.join(GOLDEN, "fast5", "rna_read_100_ch_328.fast5"), 500, 440, 6)
    lens[1] = 203
    x[1, 203:] = 0
    lo = O.seq_len_out(lens, 500 / cfg.out_len(500))
    ref = O.inference(x, lo, cfg, t)
    got = TorchCpuModel(cfg, t).inference(x, lo)
    assert np.abs(got - ref).max() < 2e-3
    assert O.ctc_greedy(got, lo) == O.ctc_greedy(ref, lo)
