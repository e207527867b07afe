"""RNA_default (k=13 / stride-5 first block, per-direction MultiRNNCell stack, T = L/5) on the GPU vs the oracle.

The reference ships no golden outputs for RNA: the oracle's RNA path is "parity unpinned", so this test
pins the CUDA path to the oracle only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from chiron_b200 import chiron_input, fast5
from oracle import chiron_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("tc", 5e-3)])
def test_rna_default_matches_oracle(rna_model, precision, tol):
    from chiron_b200.engine import Basecaller
    cfg, t, _ = rna_model
    assert cfg.k[:3] == [13, 3, 3] and cfg.stride[:3] == [5, 1, 1] and cfg.rnn_layout == 1
    sig = fast5.read_raw_signal(os.path.join(GOLDEN, "fast5", "rna_read_100_ch_328.fast5"))[::-1].astype(np.float32)
    L, jump = 500, 440
    x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), L, jump)
    T = cfg.out_len(L)
    assert T == 100
    lens_o = O.seq_len_out(lens, L / T)
    ref = O.inference(x, lens_o, cfg, t)
    bc = Basecaller("RNA_default", device=0, precision=precision)
    assert bc.out_len(L) == T and bc.out_len(2000) == 400
    bases, n_bases, prob, logits = bc.basecall_batch(x, lens, beam=0, want_logits=True)
    assert np.abs(logits - ref).max() < tol
    got = [bases[b, :n_bases[b]].tolist() for b in range(len(x))]
    assert got == O.ctc_greedy(ref, lens_o)
    # BASELINE config 3: beam search width 50 on the same logits, and the `simple` assembly kernel (440 <= 0.9*500)
    b50, n50, _, _ = bc.basecall_batch(x, lens, beam=50)
    ref50 = O.ctc_decode_c(ref, lens_o, 50)
    mism = sum(b50[b, :n50[b]].tolist() != ref50[b] for b in range(len(x)))
    assert mism == 0, "%d of %d windows differ under beam search" % (mism, len(x))
    seq, qual, pos = bc.assemble(b50, n50, prob, jump, L)
    keep = [b for b in range(len(x)) if n50[b] > 0]
    segs = [O.index2base(b50[b, :n50[b]]) for b in keep]
    cons, cq, ref_pos = O.simple_assembly_qs(segs, [prob[b] for b in keep], jump / L, kernal="simple")
    assert O.get_assembler_kernal(jump, L) == "simple"
    assert seq == O.index2base(np.argmax(cons, axis=0)) and pos[keep].tolist() == ref_pos.tolist()
    covered = cons.sum(axis=0) > 0
    ref_q = O.qs_string(cons, cq)
    assert [c for c, ok in zip(qual, covered) if ok] == [c for c, ok in zip(ref_q, covered) if ok]
    bc.close()
