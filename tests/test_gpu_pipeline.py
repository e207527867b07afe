"""GPU end-to-end: `chiron call` / chiron_eval.run on the bundled reads vs the reference's golden output tree."""
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN, read_fasta_records

pytestmark = pytest.mark.gpu


def _read(path):
    with open(path) as f:
        return f.read()


@pytest.mark.parametrize("precision", [None, "fp32"], ids=["default_tc", "fp32"])
def test_chiron_call_on_signal_folder_matches_golden_tree(tmp_path, precision):
    """`chiron call -p dna-pre` (batch 400, L 400, jump 390, beam 30): result/read1.fastq and segments/read1.fastq are
    byte-identical to the reference's files; read3's segments equal the golden ones after undoing the reference's
    collation rotation (SURVEY.md finding 6) and its consensus is therefore produced from windows in true order.
    Run as a user runs it (no --precision: the tensor-core kernels) and with the FFMA kernels."""
    import json
    from chiron_b200 import entry
    out = str(tmp_path / "out")
    entry.main(["call", "-i", os.path.join(GOLDEN, "DNA", "raw"), "-o", out, "-m", "DNA_default", "-p", "dna-pre"]
               + (["--precision", precision] if precision else []))
    with open(os.path.join(out, "meta", "all.perf.json")) as f:
        assert json.load(f)["precision"] == (precision or "tc")
    assert _read(os.path.join(out, "result", "read1.fastq")) == _read(os.path.join(GOLDEN, "DNA", "result", "read1.fastq"))
    assert _read(os.path.join(out, "segments", "read1.fastq")) == _read(os.path.join(GOLDEN, "DNA", "segments", "read1.fastq"))
    segs3 = read_fasta_records(os.path.join(out, "segments", "read3.fastq"))
    gold3 = read_fasta_records(os.path.join(GOLDEN, "DNA", "segments", "read3.fastq"))
    assert segs3[112:] + segs3[:112] == gold3
    meta = _read(os.path.join(out, "meta", "read1.meta")).split("\n")
    assert meta[0] == "# Reading Basecalling assembly output total rate(bp/s)"
    assert meta[3] == "2589 400 400 390 0"
    assert os.path.exists(os.path.join(out, "meta", "all.meta"))
    assert os.path.isdir(os.path.join(out, "raw")) and os.path.isdir(os.path.join(out, "reference"))


def test_fast5_input_greedy_fasta_and_concise(tmp_path):
    """fast5 in -> same bases as the .signal path; greedy decoder; fasta output has no trailing newline (:220)."""
    from chiron_b200 import entry
    src = tmp_path / "in"
    src.mkdir()
    shutil.copy(os.path.join(GOLDEN, "fast5", "read1.fast5"), str(src / "read1.fast5"))
    out_a, out_b = str(tmp_path / "a"), str(tmp_path / "b")
    common = ["-m", "DNA_default", "-l", "300", "-j", "290", "-b", "100", "--beam", "0", "-e", "fasta"]      # default precision: tc
    entry.main(["call", "-i", str(src), "-o", out_a] + common + ["--concise"])
    entry.main(["call", "-i", os.path.join(GOLDEN, "DNA", "raw", "read1.signal"), "-o", out_b] + common)
    fa = _read(os.path.join(out_a, "result", "read1.fasta"))
    assert fa.startswith(">read1\n") and not fa.endswith("\n") and set(fa.split("\n")[1]) <= set("ACGT")
    assert fa == _read(os.path.join(out_b, "result", "read1.fasta"))
    assert not os.path.exists(os.path.join(out_a, "segments", "read1.fasta"))      # --concise
    assert os.path.exists(os.path.join(out_b, "segments", "read1.fasta"))
    assert len(fa.split("\n")[1]) > 2000


def test_batch_composition_does_not_change_results(tmp_path):
    """Windows are packed across reads; with population BatchNorm the result of a read must not depend on batch size:
    batches of 37 windows (GPU batch forced down to the flag) against the default packing (>= 4096 windows per batch)."""
    import types
    from chiron_b200 import chiron_eval
    outs = []
    for bs in (37, 400):
        if bs == 37:
            os.environ["CHIRON_B200_GPU_BATCH"] = "1"
        else:
            os.environ.pop("CHIRON_B200_GPU_BATCH", None)
        out = str(tmp_path / ("o%d" % bs))
        flags = types.SimpleNamespace(input=os.path.join(GOLDEN, "DNA", "raw"), output=out, model="DNA_default", start=0,
                                      batch_size=bs, segment_len=400, jump=390, threads=0, beam=0, extension="fastq",
                                      concise=True, mode="dna", preset=None, recursive=True, reverse_fast5=False,
                                      precision=None)
        chiron_eval.run(flags)
        outs.append(_read(os.path.join(out, "result", "read3.fastq")))
    os.environ.pop("CHIRON_B200_GPU_BATCH", None)
    assert outs[0] == outs[1]


def test_two_slot_async_pipeline_matches_synchronous_call(dna_model):
    """cb_basecall_submit / cb_basecall_collect (pinned staging, copy-in / compute / copy-out streams, two batches in
    flight) return exactly what the synchronous cb_basecall_host returns, in submission order, for batches of different
    sizes (the staging buffers and the workspace grow while a batch is in flight)."""
    import numpy as np
    from chiron_b200.engine import Basecaller
    from oracle import chiron_oracle as O
    cfg, t, _ = dna_model
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), 300, 290)
    cuts = [(0, 40), (40, 150), (150, 151), (151, len(x))]
    bc = Basecaller("DNA_default", device=0)
    want = [bc.basecall_batch(x[a:b], lens[a:b], beam=0)[:3] for a, b in cuts]
    got, inflight = [], []
    for i, (a, b) in enumerate(cuts):
        if len(inflight) == 2:
            got.append(bc.basecall_collect(inflight.pop(0)))
        inflight.append(bc.basecall_submit(i & 1, x[a:b], lens[a:b], beam=0))
    while inflight:
        got.append(bc.basecall_collect(inflight.pop(0)))
    for (wb, wn, wp), (gb, gn, gp) in zip(want, got):
        assert np.array_equal(wb, gb) and np.array_equal(wn, gn) and np.array_equal(wp, gp)
    with pytest.raises(Exception):
        bc.basecall_collect((0, 1, bc.out_len(300)))          # empty slot
    bc.close()


@pytest.mark.parametrize("jump,kernel,oracle_reads", [(270, "simple", (1, 3)), (290, "glue", (1, 2, 3, 4, 5)), (300, "stick", (1, 3))])
def test_config1_five_reads_three_assembly_regimes(tmp_path, dna_model, jump, kernel, oracle_reads):
    """BASELINE config 1 as stated: all five bundled reads (1,046,731 samples), `-l 300 -b 100 --beam 0`, in the default
    (tensor-core) precision, at the three `-j` regimes that select the three assembly kernels (270 simple,
    290 glue, 300 stick).  For every read: window coordinates, consensus and quality string equal the oracle's
    simple_assembly of the decoded segments (the reference's easy_assembler restated, pinned by tests/golden/assembly_ref);
    for the reads in ``oracle_reads`` (all five at -j 290) every window's greedy bases equal the CPU oracle's."""
    import types
    from chiron_b200 import chiron_eval
    from oracle import chiron_oracle as O
    cfg, t, _ = dna_model
    L = 300
    assert O.get_assembler_kernal(jump, L) == kernel
    out = str(tmp_path / "out")
    flags = types.SimpleNamespace(input=os.path.join(GOLDEN, "DNA", "raw"), output=out, model="DNA_default", start=0,
                                  batch_size=100, segment_len=L, jump=jump, threads=0, beam=0, extension="fastq",
                                  concise=False, mode="dna", preset=None, recursive=False, reverse_fast5=False, precision=None)
    os.makedirs(out)
    summary = chiron_eval.evaluation(flags)
    assert flags.precision_used == "tc" and len(summary) == 5
    assert sum(v["samples"] for v in summary.values()) == 1046731
    for n in range(1, 6):
        name = "read%d" % n
        segs = read_fasta_records(os.path.join(out, "segments", name + ".fastq"))
        info = summary[name + ".signal"]
        pos, kept = np.asarray(info["pos"]), np.asarray(info["kept"])      # coordinates may be negative (simple kernel)
        assert len(segs) == int(kept.sum()) and (pos[~kept] == -1).all()
        result = _read(os.path.join(out, "result", name + ".fastq")).split("\n")
        if n in oracle_reads:                         # per-window bases against the oracle's forward pass + greedy decoder
            sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", name + ".signal"))
            ref = O.basecall_signal(sig, cfg, t, L, jump, beam=0, batch=512)
            assert segs == ref["segments"], "%s: windows decode differently from the oracle" % name
            assert pos[kept].tolist() == ref["pos"].tolist()
            assert result[1] == ref["consensus"] and result[3] == ref["qual"]
        else:                                         # coordinates / consensus from the decoded segments themselves
            cons, _, ref_pos = O.simple_assembly_qs(segs, None, jump / L, kernal=kernel)
            assert pos[kept].tolist() == ref_pos.tolist()
            assert result[1] == (O.index2base(np.argmax(cons, axis=0)) if cons.shape[1] else "")
        assert result[0] == "@" + name and len(result[3]) == len(result[1])
