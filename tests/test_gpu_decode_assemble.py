"""GPU parity for the beam-search decoder and the overlap assembly, against the reference's golden files and the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, read_fasta_records
from oracle import chiron_oracle as O

pytestmark = pytest.mark.gpu
L, JUMP, BEAM = 400, 390, 30
B2I = {"A": 0, "C": 1, "G": 2, "T": 3}


@pytest.fixture(scope="module", params=["tc", "fp32"])
def caller(request):
    """Both arithmetic modes: "tc" is what `chiron call` and bench.py run, "fp32" the FFMA kernels."""
    from chiron_b200.engine import Basecaller
    bc = Basecaller("DNA_default", device=0, precision=request.param)
    yield bc
    bc.close()


def _pack(segs, T):
    bases = np.zeros((len(segs), T), dtype=np.int8)
    n = np.zeros(len(segs), dtype=np.int32)
    for i, s in enumerate(segs):
        bases[i, :len(s)] = [B2I[c] for c in s]
        n[i] = len(s)
    return bases, n


def _result(name):
    with open(os.path.join(GOLDEN, "DNA", "result", name + ".fastq")) as f:
        lines = f.read().split("\n")
    return lines[1], lines[3]


def test_read1_signal_to_fastq_is_byte_exact(caller, dna_model):
    """raw/read1.signal -> [GPU: forward, beam 30, glue assembly, qs] == segments/read1.fastq and result/read1.fastq."""
    cfg, _, _ = dna_model
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), L, JUMP)
    bases, n_bases, prob, _ = caller.basecall_batch(x, lens, beam=BEAM)
    segs = [O.index2base(bases[b, :n_bases[b]]) for b in range(len(x)) if n_bases[b] > 0]
    assert segs == read_fasta_records(os.path.join(GOLDEN, "DNA", "segments", "read1.fastq"))
    seq, qual, pos = caller.assemble(bases, n_bases, prob, JUMP, L)
    gold_seq, gold_qual = _result("read1")
    assert seq == gold_seq
    assert qual == gold_qual
    assert pos[0] == 0 and (np.diff(pos) >= 0).all()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_assembly_of_golden_segments(caller, n):
    segs = read_fasta_records(os.path.join(GOLDEN, "DNA", "segments", "read%d.fastq" % n))
    bases, nb = _pack(segs, 64)
    seq, _, pos = caller.assemble(bases, nb, None, JUMP, L, with_qs=False)
    assert seq == _result("read%d" % n)[0]
    _, _, ref_pos = O.simple_assembly_qs(segs, None, JUMP / L, kernal="glue")
    assert pos.tolist() == ref_pos.tolist()                       # simple_assembly coordinates, bit exact


@pytest.mark.parametrize("kernel,jump,seglen", [("simple", 200, 400), ("simple", 440, 500), ("stick", 300, 300),
                                                ("glue", 290, 300)])
def test_assembly_kernels_vs_oracle(caller, kernel, jump, seglen):
    rng = np.random.default_rng(11)
    genome = "".join(rng.choice(list("ACGT"), size=4000))
    segs, p = [], 0
    while p < len(genome) - 60:
        ln = int(rng.integers(20, 60))
        s = list(genome[p:p + ln])
        for _ in range(int(rng.integers(0, 4))):                 # sprinkle basecalling errors
            s[int(rng.integers(0, len(s)))] = "ACGT"[int(rng.integers(0, 4))]
        segs.append("".join(s))
        p += int(rng.integers(5, ln)) if kernel == "simple" else ln - int(rng.integers(0, 3))
    segs.insert(7, "")                                            # an empty window is skipped like sparse2dense does
    qs = rng.uniform(1.0, 9.0, size=len(segs)).astype(np.float32)
    bases, nb = _pack(segs, 64)
    seq, qual, pos = caller.assemble(bases, nb, qs, jump, seglen, kernel=kernel)
    ne = [i for i, s in enumerate(segs) if s]
    cons, cq, ref_pos = O.simple_assembly_qs([segs[i] for i in ne], [qs[i] for i in ne], jump / seglen, kernal=kernel)
    assert O.get_assembler_kernal(jump, seglen) == kernel
    assert pos[ne].tolist() == ref_pos.tolist() and pos[7] == -1
    assert seq == O.index2base(np.argmax(cons, axis=0))
    ref_q = O.qs_string(cons, cq)
    covered = cons.sum(axis=0) > 0
    assert "".join(c for c, ok in zip(qual, covered) if ok) == "".join(c for c, ok in zip(ref_q, covered) if ok)


def test_assembly_edge_cases(caller):
    bases, nb = _pack(["ACGT"], 8)
    seq, qual, pos = caller.assemble(bases, nb, np.ones(1, np.float32), 390, 400)
    assert seq == "" and qual == "" and pos.tolist() == [0]      # reference quirk: a single window never sets `length`
    bases, nb = _pack(["", ""], 8)
    seq, _, pos = caller.assemble(bases, nb, None, 390, 400, with_qs=False)
    assert seq == "" and pos.tolist() == [-1, -1]


def test_beam_kernel_matches_oracle_on_synthetic_logits(caller):
    import torch
    rng = np.random.default_rng(5)
    B, T = 96, 120
    lg = rng.normal(scale=2.5, size=(B, T, 5)).astype(np.float32)
    lg[:, :, 4] += 3.0
    lg[:, ::5, :] = np.round(lg[:, ::5, :])
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0], lens[1] = T, 0
    for W in (1, 8, 50):
        ref = O.ctc_decode_c(lg, lens, W)
        bases, nb = caller.decode_device(torch.from_numpy(lg).cuda(), torch.from_numpy(lens).cuda(), beam=W)
        bases, nb = bases.cpu().numpy(), nb.cpu().numpy()
        got = [bases[b, :nb[b]].tolist() for b in range(B)]
        mism = sum(g != r for g, r in zip(got, ref))
        # device expf/log1pf differ from glibc by <= 2 ulp; beam pruning can amplify that on adversarial random logits
        assert mism <= 2, "%d/%d windows differ at beam width %d" % (mism, B, W)


def test_cooperative_beam_kernel_is_bit_identical_to_the_sequential_one(caller):
    """beam_warp_kernel (a warp walks one window cooperatively: parallel rank sort, per-branch probability updates, warp
    reduction for the beam bottom, pre-filtered extension loop) against beam_kernel (one thread runs cb_beam_decode_one, the
    routine the CPU tests pin to the oracle): same device libm, so the outputs must be IDENTICAL -- on tie-heavy random
    logits, ragged lengths, widths from 1 to 100."""
    import torch
    rng = np.random.default_rng(11)
    B, T = 200, 150
    lg = rng.normal(scale=2.0, size=(B, T, 5)).astype(np.float32)
    lg[:, :, 4] += 2.0
    lg[:, ::3, :] = np.round(lg[:, ::3, :])            # exact ties
    lg[50:60] = 0.0                                    # all candidates equal
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[:4] = [T, 0, 1, 2]
    dl, dn = torch.from_numpy(lg).cuda(), torch.from_numpy(lens).cuda()
    old = {k: os.environ.get(k) for k in ("CB_BEAM_SMEM", "CB_BEAM_STAGE_LOGITS")}
    try:
        for W in (1, 2, 3, 30, 50, 100):
            out = {}
            # the thread-per-window fallback, and the cooperative kernel with the logits read from global memory / staged in
            # shared memory (the launcher picks between the last two by occupancy; here each is forced)
            for mode, (smem, stage) in {"fallback": ("0", "0"), "global": ("1", "0"), "staged": ("1", "1")}.items():
                os.environ["CB_BEAM_SMEM"], os.environ["CB_BEAM_STAGE_LOGITS"] = smem, stage
                bases, nb = caller.decode_device(dl, dn, beam=W)
                torch.cuda.synchronize()
                out[mode] = (bases.cpu().numpy().copy(), nb.cpu().numpy().copy())
            for mode in ("global", "staged"):
                assert np.array_equal(out["fallback"][1], out[mode][1]), "n_bases differ at width %d (%s)" % (W, mode)
                assert np.array_equal(out["fallback"][0], out[mode][0]), "bases differ at width %d (%s)" % (W, mode)
            assert out["global"][1].sum() > 0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_serving_signature_predict(caller, dna_model):
    """{x, seq_len} -> {indices, values, dense_shape, logits, prob_logits, log_prob} (chiron/export_test.py:103-113): the
    sparse triple is the golden read1 segments, logits / prob_logits the oracle's, and log_prob the score the oracle's
    restatement of TopPaths() gives on the SAME logits (device expf/log1pf vs glibc: a few ulp over 400 frames)."""
    cfg, t, _ = dna_model
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), L, JUMP)
    x, lens = x[-24:], lens[-24:]                         # incl. the ragged last window
    out = caller.predict(x, lens, beam_width=BEAM)
    N, T = len(x), out["logits"].shape[1]
    assert out["logits"].shape == (N, T, 5) and out["prob_logits"].shape == (N,) and out["log_prob"].shape == (N, 1)
    ref = O.inference(x, lens, cfg, t)
    assert np.abs(out["logits"] - ref).max() < 5e-3
    assert np.allclose(out["prob_logits"], O.path_prob(ref), atol=1e-3)
    len_out = O.seq_len_out(lens, L / T)
    paths, scores = O.ctc_beam_scores_c(out["logits"], len_out, BEAM)
    golden = read_fasta_records(os.path.join(GOLDEN, "DNA", "segments", "read1.fastq"))[-N:]
    assert [O.index2base(p) for p in paths] == golden
    rows = out["indices"][:, 0]
    got = [out["values"][rows == b].tolist() for b in range(N)]
    assert got == paths
    assert all((out["indices"][rows == b, 1] == np.arange((rows == b).sum())).all() for b in range(N))
    assert out["dense_shape"].tolist() == [N, max(len(p) for p in paths)]
    assert out["indices"].dtype == np.int64 and out["values"].dtype == np.int64
    assert np.allclose(out["log_prob"][:, 0], scores, rtol=0, atol=2e-3), np.abs(out["log_prob"][:, 0] - scores).max()
