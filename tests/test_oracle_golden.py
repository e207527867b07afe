"""Pin the CPU oracle against the reference's own golden artefacts.

Fixtures are verbatim copies of chiron/example_data/DNA/output/{raw,segments,result} (produced by the reference with
`-b 400 -l 400 -j 390 --beam 30`, meta/read1.meta:3-4)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, read_fasta_records
from oracle import chiron_oracle as O

L, JUMP, BEAM = 400, 390, 30


def _segments(name):
    return read_fasta_records(os.path.join(GOLDEN, "DNA", "segments", name + ".fastq"))


def _result(name):
    with open(os.path.join(GOLDEN, "DNA", "result", name + ".fastq")) as f:
        lines = f.read().split("\n")
    return lines[1], lines[3]


@pytest.fixture(scope="module")
def read1(dna_model):
    cfg, t, _ = dna_model
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    return O.basecall_signal(sig, cfg, t, L, JUMP, beam=0)      # greedy paths; logits reused below


def test_read1_end_to_end_beam30(read1):
    """raw/read1.signal -> segments/read1.fastq: 161/161 windows; result/read1.fastq sequence and quality string."""
    paths = O.ctc_decode_c(read1["logits"], read1["lens"], BEAM)
    segs = [O.index2base(p) for p in paths if len(p)]
    gold = _segments("read1")
    assert len(gold) == 161
    assert segs == gold
    keep = [i for i, p in enumerate(paths) if len(p)]
    cons, cq, _ = O.simple_assembly_qs(segs, [read1["path_prob"][i] for i in keep], JUMP / L, kernal="glue")
    seq, qual = _result("read1")
    assert O.index2base(np.argmax(cons, axis=0)) == seq
    assert O.qs_string(cons, cq) == qual


def test_python_and_c_beam_agree(read1):
    lg, ln = read1["logits"][:6], read1["lens"][:6]
    assert O.ctc_beam_search(lg, ln, BEAM) == O.ctc_decode_c(lg, ln, BEAM)
    assert O.ctc_greedy(read1["logits"], read1["lens"]) == O.ctc_decode_c(read1["logits"], read1["lens"], 0)


def test_read3_end_to_end_beam30(dna_model):
    """raw/read3.signal -> segments/read3.fastq: 319 non-empty windows.  The golden file is rotated by the
    reference's batch-position collation bug (chiron_eval.py:413-428,436): windows [112..318]+[0..111]."""
    cfg, t, _ = dna_model
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read3.signal"))
    x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), L, JUMP)
    assert lens[-1] == 6 and lens[-2] == 396                    # partial windows incl. the 6-sample one
    logits = np.concatenate([O.inference(x[i:i + 80], lens[i:i + 80], cfg, t) for i in range(0, len(x), 80)])
    paths = O.ctc_decode_c(logits, lens, BEAM)
    assert len(paths[-1]) == 0                                  # the 6-sample window decodes empty and is dropped
    segs = [O.index2base(p) for p in paths if len(p)]
    gold = _segments("read3")
    assert len(gold) == 319 and len(segs) == 319
    assert segs[112:] + segs[:112] == gold


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_assembly_goldens(n):
    """segments/readN.fastq -> result/readN.fastq through the glue kernel at js_ratio 390/400."""
    segs = _segments("read%d" % n)
    assert O.get_assembler_kernal(JUMP, L) == "glue"
    cons, _, pos = O.simple_assembly_qs(segs, None, JUMP / L, kernal="glue")
    assert O.index2base(np.argmax(cons, axis=0)) == _result("read%d" % n)[0]
    assert len(pos) == len(segs) and pos[0] == 0 and (np.diff(pos) >= 0).all()


def test_kernel_selection():
    assert O.get_assembler_kernal(270, 300) == "simple"
    assert O.get_assembler_kernal(290, 300) == "glue"
    assert O.get_assembler_kernal(300, 300) == "stick"
    assert O.get_assembler_kernal(440, 500) == "simple"        # BASELINE config 3


def test_windows_and_seq_len():
    sig = np.arange(1000, dtype=np.float32)
    x, lens = O.make_windows(sig, 300, 290)
    assert x.shape == (4, 300) and lens.tolist() == [300, 300, 300, 130]
    assert x[3, 129] == 999 and x[3, 130] == 0
    assert O.seq_len_out(np.array([500, 12, 13, 2, 3]), 5.0).tolist() == [100, 2, 3, 0, 1]   # half to even


def test_fp32_vs_fp64_noise_floor(dna_model):
    cfg, t, _ = dna_model
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    x, lens = O.make_windows(O.normalize_signal(sig, 1), L, JUMP)
    a = O.inference(x[:8], lens[:8], cfg, t, np.float32)
    b = O.inference(x[:8], lens[:8], cfg, t, np.float64)
    assert np.abs(a - b).max() < 2e-3
    assert (a.argmax(2) == b.argmax(2)).all()
