"""CPU-side tests: the C-ABI library loads and exports every declared symbol, and the host-compiled instantiations of
the kernels' sequential routines (beam search, assembly displacements) agree with the oracle.  No GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN, ROOT, read_fasta_records
from chiron_b200 import _lib
from oracle import chiron_oracle as O


def _decl_names(header):
    with open(os.path.join(ROOT, "include", header)) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _decl_names("chiron_b200.h")
    assert "cb_forward" in names and "cb_assemble" in names and len(names) >= 20
    for n in names + _decl_names("chiron_b200_selftest.h"):
        assert hasattr(lib, n), "libchiron_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.cb_version().startswith(b"chiron_b200")


def test_bad_blob_is_rejected_without_a_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    junk = ctypes.create_string_buffer(b"NOPE" + bytes(400))
    rc = lib.cb_create(ctypes.cast(junk, ctypes.c_void_p), 404, 0, 0, ctypes.byref(h))
    assert rc == -2 and b"CBW1" in lib.cb_last_error()
    with pytest.raises(_lib.ChironB200Error):
        _lib.check(rc, "cb_create")


def _beam(logits, length, W, pool=None):
    lib = _lib.load()
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    out = np.zeros(max(len(lg), 1), dtype=np.int8)
    pool = pool or (4 * W + 4096)
    n = lib.cb_selftest_beam(lg.ctypes.data_as(ctypes.c_void_p), int(length), lg.shape[1], W, pool,
                             out.ctypes.data_as(ctypes.c_void_p))
    assert n >= 0, n
    return out[:n].tolist()


def _disp(cur, prev, kernel, jump, L):
    lib = _lib.load()
    m = {"A": 0, "C": 1, "G": 2, "T": 3}
    a = np.array([m[c] for c in cur], dtype=np.int8)
    b = np.array([m[c] for c in prev], dtype=np.int8)
    return lib.cb_selftest_disp(a.ctypes.data_as(ctypes.c_void_p), len(a), b.ctypes.data_as(ctypes.c_void_p), len(b),
                                _lib.ASM_KERNELS[kernel], jump, L)


@pytest.fixture(scope="module")
def read1_logits(dna_model):
    cfg, t, _ = dna_model
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    x, lens = O.make_windows(O.normalize_signal(sig, 1), 400, 390)
    return O.inference(x[:24], lens[:24], cfg, t), lens[:24]


def test_beam_routine_matches_oracle_and_golden(read1_logits):
    logits, lens = read1_logits
    gold = read_fasta_records(os.path.join(GOLDEN, "DNA", "segments", "read1.fastq"))
    ref = O.ctc_decode_c(logits, lens, 30)
    for b in range(len(logits)):
        got = _beam(logits[b], lens[b], 30)
        assert got == ref[b]
        assert O.index2base(got) == gold[b]
    for W in (1, 2, 7, 50, 100):
        assert _beam(logits[3], 400, W) == O.ctc_decode_c(logits[3:4], lens[3:4], W)[0]


def test_beam_pool_compaction_is_transparent(read1_logits):
    logits, lens = read1_logits
    ref = O.ctc_decode_c(logits[:4], lens[:4], 30)
    for b in range(4):
        assert _beam(logits[b], lens[b], 30, pool=400) == ref[b]               # forces many in-place compactions


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 12), st.integers(1, 40), st.booleans())
def test_beam_routine_random_logits(seed, W, T, ties):
    rng = np.random.default_rng(seed)
    lg = rng.normal(scale=3.0, size=(T, 5)).astype(np.float32)
    if ties:
        lg = np.round(lg)
    ref = O.ctc_decode_c(lg[None], np.array([T], np.int32), W)[0]
    assert _beam(lg, T, W, pool=2 * W * (T + 1) + 2) == ref          # smallest pool that can never overflow
    assert _beam(lg, T, W) == ref


_seq = st.text(alphabet="ACGT", min_size=1, max_size=60)


@settings(max_examples=300, deadline=None)
@given(_seq, _seq)
def test_displacement_kernels_match_reference_semantics(cur, prev):
    assert _disp(cur, prev, "stick", 300, 300) == O.stick_kernal(cur, prev)
    assert _disp(cur, prev, "glue", 390, 400) == O.glue_kernal(cur, prev)
    assert _disp(cur, prev, "simple", 440, 500) == O.simple_assembly_kernal(cur, prev, 0.2, 440 / 500)


def test_simple_kernel_on_overlapping_real_segments_and_autojunk():
    segs = read_fasta_records(os.path.join(GOLDEN, "DNA", "segments", "read4.fastq"))
    for i in range(1, 200):
        # emulate a heavily overlapping jump: suffix of the previous segment + the next one
        prev, cur = segs[i - 1], segs[i - 1][len(segs[i - 1]) // 2:] + segs[i][:10]
        assert _disp(cur, prev, "simple", 200, 400) == O.simple_assembly_kernal(cur, prev, 0.2, 0.5)
        assert _disp(segs[i], segs[i - 1], "glue", 390, 400) == O.glue_kernal(segs[i], segs[i - 1])
    rng = np.random.default_rng(5)
    for n in (199, 200, 201, 260, 400):       # difflib autojunk switches on at len(b) >= 200
        prev = "".join(rng.choice(list("ACGT"), size=n))
        cur = prev[n // 3:] + "".join(rng.choice(list("ACGT"), size=n // 3))
        assert _disp(cur, prev, "simple", 250, 500) == O.simple_assembly_kernal(cur, prev, 0.2, 0.5)
        skew = "A" * (n - 3) + "CGT"          # rare letters survive the popularity purge
        assert _disp(skew[5:] + "AC", skew, "simple", 250, 500) == O.simple_assembly_kernal(skew[5:] + "AC", skew, 0.2, 0.5)
