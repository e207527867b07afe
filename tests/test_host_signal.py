"""The native host-side signal preparation (cb_host_parse_signal / cb_host_normalize / cb_host_windows,
chiron_b200/csrc/cb_host_signal.cu -- rows a1-a2 of the path: chiron/chiron_input.py:527-555, 253-292, 681-692) against
the oracle's numpy formulation: bit-identical on the reference's bundled signals and on generated inputs."""
import glob
import os
import threading
import time

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN
from chiron_b200 import _lib, chiron_input, fast5
from chiron_b200.model import NORM_FULL_MAD, NORM_NONE, NORM_UNIQUE_MAD
from oracle import chiron_oracle as O

SIGNALS = sorted(glob.glob(os.path.join(GOLDEN, "DNA", "raw", "*.signal")))


def test_bundled_signals_parse_normalise_window_like_the_oracle():
    assert len(SIGNALS) >= 2
    for path in SIGNALS:
        ref = O.read_signal_text(path)
        got = chiron_input.read_signal(path)
        assert got.dtype == np.float32 and np.array_equal(got, ref)
        for mode in (NORM_NONE, NORM_UNIQUE_MAD, NORM_FULL_MAD):
            assert np.array_equal(chiron_input.normalize_signal(got, mode), O.normalize_signal(ref, mode))
        norm = O.normalize_signal(ref, NORM_UNIQUE_MAD)
        for L, jump in ((400, 390), (300, 290), (512, 512), (500, 1000), (7, 3)):
            ds = chiron_input.windows_from_signal(norm, jump, L)
            x, lens = O.make_windows(norm, L, jump)
            assert np.array_equal(ds.event, x) and np.array_equal(ds.event_length, lens)


def test_fast5_signals_normalise_like_the_oracle():
    for name, mode in (("read1.fast5", NORM_UNIQUE_MAD), ("rna_read_100_ch_328.fast5", NORM_FULL_MAD)):
        raw = fast5.read_raw_signal(os.path.join(GOLDEN, "fast5", name)).astype(np.float32)
        for m in (mode, NORM_UNIQUE_MAD, NORM_FULL_MAD):
            assert np.array_equal(chiron_input.normalize_signal(raw, m), O.normalize_signal(raw, m))
            assert np.array_equal(chiron_input.normalize_signal(raw[::-1], m), O.normalize_signal(raw[::-1], m))


_number = st.one_of(
    st.integers(-40000, 40000).map(str),
    st.integers(-10**14, 10**14).map(str),
    st.floats(allow_nan=False, allow_infinity=False, width=64).map(repr),
    st.floats(-1e4, 1e4, allow_nan=False).map(lambda v: "%.3f" % v),
    st.sampled_from(["+7", ".5", "5.", "1E-3", "1e5", "-0", "007", "16777217", "33554433.0", "0.1", "12345678901234567890"]))
_sep = st.sampled_from([" ", "\n", "\t", "  ", " \r\n", "\n\n"])


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(_number, _sep), max_size=60), _sep)
def test_parser_matches_numpy_string_conversion(tokens, lead):
    text = lead + "".join(t + s for t, s in tokens)
    with np.errstate(over="ignore"):
        ref = np.asarray(text.split(), dtype=np.float32)
    got = chiron_input.parse_signal_text(text.encode())
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))          # bit for bit, including -0.0


def test_parser_rejects_malformed_tokens():
    """Tokens numpy cannot convert either, and (a documented limit of the native parser) tokens over 127 characters."""
    for bad in (b"1 2 x3", b"4 5.5.5", b"--3", b"1e", b"3 , 4", b"1" * 200):
        with pytest.raises(_lib.ChironB200Error):
            chiron_input.parse_signal_text(bad)
    assert chiron_input.parse_signal_text(b"").size == 0 and chiron_input.parse_signal_text(b" \n\t ").size == 0


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 400), st.integers(0, 2**31 - 1), st.sampled_from(["dac", "small_int", "float", "wide_int", "const"]))
def test_normalisation_matches_oracle_on_generated_signals(n, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "dac":                         # raw DAC integers: the presence-bitmap path
        s = rng.integers(200, 900, size=n).astype(np.float32)
    elif kind == "small_int":                 # few distinct values: even / odd unique counts, ties around the median
        s = rng.integers(-3, 4, size=n).astype(np.float32)
    elif kind == "wide_int":                  # integers beyond the bitmap range: the sort path
        s = rng.integers(-2**22, 2**22, size=n).astype(np.float32)
    elif kind == "const":
        s = np.full(n, 417, dtype=np.float32)
    else:
        s = rng.normal(0, 50, size=n).astype(np.float32)
    for mode in (NORM_UNIQUE_MAD, NORM_FULL_MAD):
        with np.errstate(all="ignore"):
            ref = O.normalize_signal(s, mode)
        got = chiron_input.normalize_signal(s, mode)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))      # incl. the inf / nan of a zero MAD


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 300), st.integers(1, 50), st.integers(1, 60), st.integers(0, 2**31 - 1))
def test_windows_match_oracle_on_generated_signals(n, jump, L, seed):
    s = np.random.default_rng(seed).normal(size=n).astype(np.float32)
    ds = chiron_input.windows_from_signal(s, jump, L)
    x, lens = O.make_windows(s, L, jump)
    assert np.array_equal(ds.event, x) and np.array_equal(ds.event_length, lens)


def test_reader_threads_run_in_parallel():
    """The native helpers release the GIL: four threads preparing reads take well under four times one thread's time
    (the reference's pure-Python token loop cannot overlap)."""
    path = SIGNALS[-1]
    data = open(path, "rb").read() * 4

    def work():
        for _ in range(6):
            s = chiron_input.parse_signal_text(data)
            chiron_input.windows_from_signal(chiron_input.normalize_signal(s, NORM_UNIQUE_MAD), 390, 400)

    work()
    t0 = time.perf_counter()
    work()
    one = time.perf_counter() - t0
    if (os.cpu_count() or 1) < 4:
        pytest.skip("needs 4 cores")
    threads = [threading.Thread(target=work) for _ in range(4)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    four = time.perf_counter() - t0
    assert four < 2.5 * one, "4 threads took %.1f ms, one thread %.1f ms" % (four * 1e3, one * 1e3)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 40), st.integers(1, 30), st.integers(0, 2**31 - 1), st.sampled_from(["read1", "r", "a/b_c.d-é"]))
def test_segment_records_match_the_reference_format(n, T, seed, name):
    """cb_host_format_segments == write_output's loop over index2base strings (chiron_eval.py:211-214)."""
    from chiron_b200.engine import format_segments, index2base
    rng = np.random.default_rng(seed)
    bases = rng.integers(0, 4, size=(n, T)).astype(np.int8)
    n_bases = rng.integers(0, T + 1, size=n).astype(np.int32)
    kept = [index2base(bases[i, :n_bases[i]]) for i in range(n) if n_bases[i] > 0]      # sparse2dense drops empty rows
    ref = "".join(">{}{}\n{}\n".format(name, str(i), r) for i, r in enumerate(kept)).encode("utf-8")
    assert format_segments(name, bases, n_bases) == ref


def test_nan_and_inf_samples_normalise_like_numpy():
    base = np.array([3, 7, 7, 1, 9, 4], dtype=np.float32)
    for bad in (np.nan, np.inf, -np.inf):
        s = base.copy()
        s[2] = bad
        for mode in (NORM_UNIQUE_MAD, NORM_FULL_MAD):
            with np.errstate(all="ignore"):
                ref = O.normalize_signal(s, mode)
            got = chiron_input.normalize_signal(s, mode)
            assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)])
    assert np.isnan(chiron_input.parse_signal_text(b"1 nan 3")[1])
