"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical inputs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import chiron_oracle as O

pytestmark = pytest.mark.gpu

# Tolerance on CTC logits (|logit| <= ~20) against the oracle; see DESIGN.md "Numerics".  The float32 oracle itself moves by
# 5.5e-4 against the float64 one on 38,400 frames and by 1.2e-3 on 32,768 others (the three LSTM layers amplify rounding
# noise ~100x, with a heavy tail: max / rms ~ 350).  "fp32" = FFMA kernels with round-to-nearest accumulation (measured
# maxima against the float64 oracle: 9.7e-4 on 38,400 frames of read1, 4.1e-3 on 262,144 frames of the bench batch).
# "tc" = tcgen05 fp16 hi/lo split with short-K partial sums and truncation compensation (1.0e-3 and 6.1e-3 on the same
# frames; rms 1.8e-5 against 1.1e-5; the round-1 kernels: 1.8e-2 and 1.1e-4 rms).  Tests on <= 40,000 frames use LOGIT_TOLS;
# the 262,144-frame check of the bench batch uses LOGIT_TOLS_LARGE.
LOGIT_TOLS = {"fp32": 2e-3, "tc": 5e-3}
LOGIT_TOLS_LARGE = {"fp32": 6e-3, "tc": 8e-3}


def _read1_windows(cfg, L=400, jump=390):
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    return O.make_windows(O.normalize_signal(sig, cfg.sig_norm), L, jump)


@pytest.fixture(scope="module", params=["fp32", "tc"])
def caller(request):
    from chiron_b200.engine import Basecaller
    bc = Basecaller("DNA_default", device=0, precision=request.param)
    bc.logit_tol = LOGIT_TOLS[request.param]
    yield bc
    bc.close()


def test_logits_and_greedy_match_oracle_read1(caller, dna_model):
    cfg, t, _ = dna_model
    x, lens = _read1_windows(cfg)
    n0 = len(x)
    x, lens = x[n0 - 48:], lens[n0 - 48:]          # includes the ragged last window
    ref = O.inference(x, lens, cfg, t)
    bases, n_bases, prob, logits = caller.basecall_batch(x, lens, beam=0, want_logits=True)
    assert np.abs(logits - ref).max() < caller.logit_tol
    ref_paths = O.ctc_greedy(ref, lens)
    got = [bases[b, :n_bases[b]].tolist() for b in range(len(x))]
    assert got == ref_paths
    np.testing.assert_allclose(prob, O.path_prob(ref), rtol=0, atol=1e-4 if caller.precision == "fp32" else 1e-3)


def test_intermediates_match_oracle(caller, dna_model):
    cfg, t, _ = dna_model
    x, lens = _read1_windows(cfg)
    x, lens = x[:8].copy(), lens[:8].copy()
    lens[3] = 57
    x[3, 57:] = 0
    caller.basecall_batch(x, lens, beam=0)
    fea = O.cnn_forward(x, cfg, t)
    got = caller.debug_fetch(0, fea.size).reshape(fea.shape)
    assert np.abs(got - fea).max() < 1e-3 * max(1.0, np.abs(fea).max())
    lasth = O.rnn_forward(fea, lens, cfg, t)
    got = caller.debug_fetch(cfg.n_layers, lasth.size).reshape(lasth.shape)
    assert np.abs(got - lasth).max() < (1e-3 if caller.precision == "fp32" else 1e-2)
    assert (got[3, 57:] == 0).all()               # dynamic_rnn: zero output past sequence_length


def test_ragged_lengths_and_tiny_batches(caller, dna_model):
    cfg, t, _ = dna_model
    x, _ = _read1_windows(cfg, L=300, jump=290)
    x = x[:5].copy()
    lens = np.array([300, 1, 17, 299, 128], dtype=np.int32)
    for b in range(5):
        x[b, lens[b]:] = 0
    ref = O.inference(x, lens, cfg, t)
    bases, n_bases, prob, logits = caller.basecall_batch(x, lens, beam=0, want_logits=True)
    assert np.abs(logits - ref).max() < caller.logit_tol
    assert [bases[b, :n_bases[b]].tolist() for b in range(5)] == O.ctc_greedy(ref, lens)
    b1 = caller.basecall_batch(x[:1], lens[:1], want_logits=True)      # B = 1
    assert np.abs(b1[3] - ref[:1]).max() < caller.logit_tol


def test_greedy_kernel_exact_on_synthetic_logits(caller):
    """Decoder in isolation (seeded synthetic logits with exact ties exercising "first maximum wins"): bit-exact
    against the C oracle, including len = 0 and len = T rows."""
    import torch
    rng = np.random.default_rng(3)
    B, T = 64, 300
    lg = rng.normal(size=(B, T, 5)).astype(np.float32)
    lg[:, :, 4] += 2.0                                 # blank-dominant like real logits
    lg[:, ::7, :] = np.round(lg[:, ::7, :])
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0] = T
    lens[1] = 0
    ref = O.ctc_decode_c(lg, lens, 0)
    dl, dn = torch.from_numpy(lg).cuda(), torch.from_numpy(lens).cuda()
    bases, n_bases = caller.decode_device(dl, dn, beam=0)
    torch.cuda.synchronize()
    bases, n_bases = bases.cpu().numpy(), n_bases.cpu().numpy()
    assert [bases[b, :n_bases[b]].tolist() for b in range(B)] == ref


def test_full_row_groups_match_oracle(caller, dna_model):
    """B > 128 with equal lengths: the first 128-window row group is full and uniform the second one is
    partial (padding rows)."""
    cfg, t, _ = dna_model
    x, lens = _read1_windows(cfg, L=200, jump=150)
    x, lens = x[:150].copy(), lens[:150].copy()
    assert (lens == 200).all()
    ref = O.inference(x, lens, cfg, t, np.float64)     # 30,000 frames: judge both modes against the float64 oracle
    bases, n_bases, prob, logits = caller.basecall_batch(x, lens, beam=0, want_logits=True)
    assert np.abs(logits - ref).max() < caller.logit_tol
    assert [bases[b, :n_bases[b]].tolist() for b in range(len(x))] == O.ctc_greedy(ref.astype(np.float32), lens)


def test_full_size_batch_properties(dna_model):
    """BASELINE size (4096 windows x 512 samples, the bench workload) through size-independent properties:
    (1) a window's result does not depend on its batch (same rows alone in a 256-window batch: bit-identical logits);
    (2) the tensor-core mode -- the mode bench.py and `chiron call` run -- decodes EXACTLY the greedy bases the fp32 FFMA
        mode decodes, for every one of the 4096 windows (the round-1 kernels differed in 9);
    (3) 512 windows sampled from the big batch agree with the oracle: logits of 262,144 frames within LOGIT_TOLS_LARGE of the
        float64 oracle (measured: tc 6.1e-3, fp32 4.1e-3), greedy bases bit-identical to the float32 oracle's in both modes."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import synthetic_windows
    from chiron_b200.engine import Basecaller
    cfg, t, _ = dna_model
    x, lens = synthetic_windows(4096, 512, 1234)
    out = {}
    for prec in ("tc", "fp32"):
        bc = Basecaller("DNA_default", device=0, precision=prec)
        bases, nb, prob, lg = bc.basecall_batch(x, lens, beam=0, want_logits=True)
        sub = slice(1000, 1256)
        b2, n2, p2, lg2 = bc.basecall_batch(x[sub], lens[sub], beam=0, want_logits=True)
        assert np.array_equal(lg[sub], lg2) and np.array_equal(bases[sub], b2) and np.array_equal(nb[sub], n2)
        out[prec] = (bases, nb, lg)
        bc.close()
    differ = [b for b in range(4096) if out["tc"][1][b] != out["fp32"][1][b] or
              not np.array_equal(out["tc"][0][b, :out["tc"][1][b]], out["fp32"][0][b, :out["fp32"][1][b]])]
    assert not differ, "tc and fp32 decode different bases in %d of 4096 windows: %s" % (len(differ), differ[:10])
    pick = np.sort(np.random.default_rng(7).choice(4096, 512, replace=False))
    ref64 = O.inference(x[pick], lens[pick], cfg, t, np.float64)
    ref_paths = O.ctc_greedy(ref64.astype(np.float32), lens[pick])
    for prec in ("tc", "fp32"):
        bases, nb, lg = out[prec]
        err = np.abs(lg[pick] - ref64).max()
        assert err < LOGIT_TOLS_LARGE[prec], "%s: max |dlogit| %.3e against the float64 oracle on 512 windows" % (prec, err)
        assert [bases[b, :nb[b]].tolist() for b in pick] == ref_paths, prec
    assert 15 < out["fp32"][1].mean() < 30          # ~20.7 bases per 512-sample window on the bundled R9 reads
