"""The fp32 path compiled for the HOST and run under a thread-per-CUDA-thread emulation (tests/cuda_emu): the same kernel
sources (cb_gemm_simt_kernel.cuh, cb_bn_kernels.cuh, cb_lstm_simt_kernel.cuh, cb_gru_simt_kernel.cuh) and the same
conv-stack orchestration (cb_conv_stack.cuh) the CUDA library is built from, checked against the oracle without a GPU.  This is what covers the batch-statistics
BatchNorm kernels and the buffer rotation of both BatchNorm modes on the CPU side; the `-m gpu` tests
cover the real launches."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from chiron_b200 import model as M
from oracle import chiron_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cuda_emu", "emu_fp32_path.cpp")
LIB = os.path.join(HERE, "cuda_emu", "_build", "libemu_fp32_path.so")
CSRC = os.path.join(os.path.dirname(HERE), "chiron_b200", "csrc")
BN_EPS = np.float32(1e-5)
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emu():
    deps = [SRC, os.path.join(HERE, "cuda_emu", "cuda_emu.h")] + [
        os.path.join(CSRC, f) for f in ("cb_simt_types.h", "cb_gemm_simt_kernel.cuh", "cb_bn_kernels.cuh", "cb_conv_stack.cuh",
                                        "cb_gru_simt_kernel.cuh", "cb_lstm_simt_kernel.cuh", "cb_stem_kernel.cuh",
                                        "cb_head_decode_kernels.cuh", "cb_seq_kernels.cuh", "cb_seq_algos.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-o", LIB, SRC])
    lib = ctypes.CDLL(LIB)
    lib.emu_conv_stack.restype = ctypes.c_int
    lib.emu_bn_stats.restype = ctypes.c_int
    lib.emu_bn_rank1.restype = ctypes.c_int
    lib.emu_gru.restype = ctypes.c_int
    lib.emu_lstm.restype = ctypes.c_int
    lib.emu_beam_small_pool.restype = ctypes.c_longlong
    return lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(FP)


def _ref_inv_shift(mean, var, scale, offset):
    inv = scale * (np.float32(1.0) / np.sqrt(var.astype(np.float32) + BN_EPS))
    return inv, offset - mean.astype(np.float32) * inv


@pytest.mark.parametrize("M_rows,C,sms", [(1, 4, 1), (37, 20, 1), (1000, 256, 1), (513, 64, 3), (300, 1024, 1), (5, 8, 148)])
def test_column_statistics_kernels(emu, M_rows, C, sms):
    rng = np.random.default_rng(M_rows + C)
    X = (rng.normal(size=(M_rows, C)) * rng.uniform(0.1, 30, size=C) + rng.normal(size=C) * 5).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, size=C).astype(np.float32)
    offset = rng.normal(size=C).astype(np.float32)
    inv, shift = np.zeros(C, np.float32), np.zeros(C, np.float32)
    assert emu.emu_bn_stats(_fp(X), ctypes.c_longlong(M_rows), C, sms, _fp(scale), _fp(offset), _fp(inv), _fp(shift)) == 0
    x64 = X.astype(np.float64)
    r_inv, r_shift = _ref_inv_shift(x64.mean(axis=0), x64.var(axis=0), scale, offset)
    np.testing.assert_allclose(inv, r_inv, rtol=2e-6, atol=0)
    np.testing.assert_allclose(shift, r_shift, rtol=0, atol=2e-6 * max(1.0, np.abs(r_shift).max()))


@pytest.mark.parametrize("B,t_in,stride,sms", [(3, 50, 1, 1), (7, 503, 5, 1), (2, 9, 4, 148), (300, 40, 2, 2)])
def test_rank1_statistics_kernels(emu, B, t_in, stride, sms):
    rng = np.random.default_rng(B * t_in)
    C = 12
    t_out = -(-t_in // stride)
    x = rng.normal(-0.2, 0.5, size=(B, t_in)).astype(np.float32)
    w = rng.normal(size=C).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, size=C).astype(np.float32)
    offset = rng.normal(size=C).astype(np.float32)
    inv, shift = np.zeros(C, np.float32), np.zeros(C, np.float32)
    assert emu.emu_bn_rank1(_fp(x), B, t_in, stride, t_out, C, sms, _fp(w), _fp(scale), _fp(offset), _fp(inv), _fp(shift)) == 0
    prod = x[:, ::stride][:, :t_out].astype(np.float64)[:, :, None] * w.astype(np.float64)
    r_inv, r_shift = _ref_inv_shift(prod.mean(axis=(0, 1)), prod.var(axis=(0, 1)), scale, offset)
    np.testing.assert_allclose(inv, r_inv, rtol=5e-6, atol=0)
    np.testing.assert_allclose(shift, r_shift, rtol=0, atol=5e-6 * max(1.0, np.abs(r_shift).max()))
    # a branch without BatchNorm: identity
    assert emu.emu_bn_rank1(_fp(x), B, t_in, stride, t_out, C, sms, _fp(w), None, None, _fp(inv), _fp(shift)) == 0
    assert (inv == 1).all() and (shift == 0).all()


def _geom(cfg):
    k = list(cfg.k) + [0] * (8 - len(cfg.k))
    s = list(cfg.stride) + [0] * (8 - len(cfg.stride))
    return (ctypes.c_int * 21)(cfg.n_blocks, cfg.channels, cfg.branch1_bn_mask, *k[:8], *s[:8], cfg.stem_k, cfg.stem_stride)


def _stem_arrays(cfg, t):
    if not cfg.stem_k:
        return []
    inv, sh = _fold(t, "conv_layer/conv1")
    return [t["conv_layer/conv1/weights"], inv, sh, t["conv_layer/conv1_bn/scale"], t["conv_layer/conv1_bn/offset"]]


def _run_stack(emu, bn_mode, cfg, ptr_arrays, rank1, x, sms, stem=()):
    B, L = x.shape
    keep = [np.ascontiguousarray(a, dtype=np.float32) if a is not None else None for a in ptr_arrays]
    tab = (FP * len(keep))(*[_fp(a) for a in keep])
    r_keep = [np.ascontiguousarray(a, dtype=np.float32) for a in rank1]
    r_tab = (FP * max(len(r_keep), 1))(*[_fp(a) for a in r_keep])
    s_keep = [np.ascontiguousarray(a, dtype=np.float32) for a in stem]
    s_tab = (FP * max(len(s_keep), 1))(*[_fp(a) for a in s_keep])
    out = np.zeros(B * L * cfg.channels, np.float32)
    n_launch = ctypes.c_longlong(0)
    T = emu.emu_conv_stack(bn_mode, _geom(cfg), tab, r_tab, s_tab, _fp(np.ascontiguousarray(x)), B, L, sms, _fp(out),
                           ctypes.byref(n_launch))
    assert T > 0, "emu_conv_stack failed (%d)" % T
    return out[:B * T * cfg.channels].reshape(B, T, cfg.channels), n_launch.value


def _fold(t, prefix):                                  # cb_create: population BN -> inv / shift
    inv = t[prefix + "_bn/scale"] * (np.float32(1.0) / np.sqrt(t[prefix + "_bn/pop_var"] + BN_EPS))
    return inv, t[prefix + "_bn/offset"] - t[prefix + "_bn/pop_mean"] * inv


TOPOLOGIES = [dict(n_blocks=3, channels=8, k=[3, 3, 3], stride=[1, 1, 1], branch1_bn_mask=1),            # DNA_default shape
              dict(n_blocks=3, channels=20, k=[13, 3, 3], stride=[5, 1, 1], branch1_bn_mask=1),         # RNA_default shape
              dict(n_blocks=4, channels=12, k=[5, 3, 7, 2], stride=[2, 1, 3, 1], branch1_bn_mask=0b0110),
              dict(n_blocks=5, channels=8, k=[3] * 5, stride=[1] * 5, branch1_bn_mask=1),             # rna_test
              dict(n_blocks=3, channels=8, k=[3] * 3, stride=[1] * 3, branch1_bn_mask=1, stem_k=9, stem_stride=5),    # RNA_model2
              dict(n_blocks=2, channels=12, k=[3] * 2, stride=[1, 2], branch1_bn_mask=1, stem_k=14, stem_stride=7)]   # RNA_model3-like


def _inputs(cfg, seed):
    rng = np.random.default_rng(seed)
    B, L = 3, 53
    x = rng.normal(-0.16, 0.43, size=(B, L)).astype(np.float32)
    x[2, 31:] = 0
    return x


@pytest.mark.parametrize("topo", TOPOLOGIES)
@pytest.mark.parametrize("sms", [1, 148])
def test_conv_stack_batch_statistics_mode(emu, topo, sms):
    cfg = M.ModelConfig(hidden=4, **topo)
    t = M.random_tensors(cfg, seed=21)
    x = _inputs(cfg, 4)
    ptrs = []
    for b in range(cfg.n_blocks):
        p = "res_layer%d" % (b + 1)
        for conv in ("branch1/conv1", "branch2/conv2a", "branch2/conv2b", "branch2/conv2c"):
            has_bn = (p + "/" + conv + "_bn/scale") in t
            ptrs += [t[p + "/" + conv + "/weights"], t[p + "/" + conv + "_bn/scale"] if has_bn else None,
                     t[p + "/" + conv + "_bn/offset"] if has_bn else None]
    got, n_launch = _run_stack(emu, 1, cfg, ptrs, [], x, sms, _stem_arrays(cfg, t))
    ref = O.cnn_forward(x, cfg, t, np.float64, bn_mode=1)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())
    assert n_launch > 10 * cfg.n_blocks


@pytest.mark.parametrize("topo", TOPOLOGIES)
def test_conv_stack_population_mode(emu, topo):
    cfg = M.ModelConfig(hidden=4, **topo)
    t = M.random_tensors(cfg, seed=22)
    x = _inputs(cfg, 5)
    C = cfg.channels
    ptrs, rank1 = [], []
    for b in range(cfg.n_blocks):
        p = "res_layer%d" % (b + 1)
        has1 = cfg.branch1_bn_mask >> b & 1
        inv1, sh1 = _fold(t, p + "/branch1/conv1") if has1 else (np.ones(C, np.float32), np.zeros(C, np.float32))
        inva, sha = _fold(t, p + "/branch2/conv2a")
        invb, shb = _fold(t, p + "/branch2/conv2b")
        invc, shc = _fold(t, p + "/branch2/conv2c")
        w1, w2a = t[p + "/branch1/conv1/weights"], t[p + "/branch2/conv2a/weights"]
        w2b, w2c = t[p + "/branch2/conv2b/weights"], t[p + "/branch2/conv2c/weights"]
        if b == 0 and not cfg.stem_k:
            rank1 = [w2a.reshape(-1), inva, sha, w1.reshape(-1), inv1, sh1]
            ptrs += [np.zeros(4, np.float32), np.zeros(4, np.float32)]                      # conv2a of block 1 is generated
            ptrs += [w2b.reshape(-1, C) * invb, shb, w2c * invc, shc]
        else:
            ptrs += [w2a * inva, sha, w2b.reshape(-1, C) * invb, shb, np.concatenate([w2c * invc, w1 * inv1]), shc + sh1]
    if not rank1:
        rank1 = [np.zeros(4, np.float32)] * 6
    got, _ = _run_stack(emu, 0, cfg, ptrs, rank1, x, 1, _stem_arrays(cfg, t))
    ref = O.cnn_forward(x, cfg, t, np.float64, bn_mode=0)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("rg,B,T,H", [(1, 5, 9, 8), (2, 37, 6, 20), (4, 70, 5, 12), (1, 3, 4, 100)])
def test_gru_recurrence_kernel(emu, rg, B, T, H):
    """gru_simt_kernel (both directions, ragged lengths incl. 0 and T, partial last CTA) against the oracle's GRUCell
    restatement fed with the same hoisted input projection."""
    rng = np.random.default_rng(rg * 1000 + B)
    D = 16
    x = rng.normal(size=(B, T, D)).astype(np.float32)
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0] = T
    if B > 1:
        lens[1] = 0
    w = {}
    for d in ("fw", "bw"):
        w[d] = dict(gk=rng.uniform(-0.4, 0.4, size=(D + H, 2 * H)).astype(np.float32),
                    gb=(rng.normal(0, 0.1, size=2 * H) + 1).astype(np.float32),
                    ck=rng.uniform(-0.4, 0.4, size=(D + H, H)).astype(np.float32),
                    cb=rng.normal(0, 0.1, size=H).astype(np.float32))
    # hoisted projection as the GEMM produces it: per direction r | u | candidate columns, biases folded in
    pre = np.zeros((B * T, 6 * H), np.float32)
    xf = x.reshape(B * T, D)
    for i, d in enumerate(("fw", "bw")):
        pre[:, i * 3 * H:i * 3 * H + 2 * H] = xf @ w[d]["gk"][:D] + w[d]["gb"]
        pre[:, i * 3 * H + 2 * H:(i + 1) * 3 * H] = xf @ w[d]["ck"][:D] + w[d]["cb"]
    out = np.full((B * T, 2 * H), -7.0, np.float32)
    rec = {d: (np.ascontiguousarray(w[d]["gk"][D:]), np.ascontiguousarray(w[d]["ck"][D:])) for d in w}
    rc = emu.emu_gru(rg, B, T, H, _fp(pre), 6 * H, _fp(rec["fw"][0]), _fp(rec["bw"][0]), _fp(rec["fw"][1]), _fp(rec["bw"][1]),
                     lens.ctypes.data_as(ctypes.c_void_p), _fp(out), 2 * H)
    assert rc == 0
    ref = np.concatenate([O.gru_direction(x, lens, w[d]["gk"], w[d]["gb"], w[d]["ck"], w[d]["cb"], d == "bw", np.float64)
                          for d in ("fw", "bw")], axis=2)
    got = out.reshape(B, T, 2 * H)
    assert np.abs(got - ref).max() < 2e-5
    for b in range(B):
        assert (got[b, lens[b]:] == 0).all()              # dynamic_rnn: zero output past sequence_length


@pytest.mark.parametrize("rg,B,T,H", [(1, 5, 9, 8), (2, 37, 6, 20), (4, 70, 5, 12), (1, 3, 4, 100)])
def test_lstm_recurrence_kernel(emu, rg, B, T, H):
    """lstm_simt_kernel (both directions, ragged lengths) against the oracle's LSTMCell fed with the same projection."""
    rng = np.random.default_rng(rg * 100 + B)
    D = 16
    x = rng.normal(size=(B, T, D)).astype(np.float32)
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0] = T
    if B > 1:
        lens[1] = 0
    kern = {d: rng.uniform(-0.4, 0.4, size=(D + H, 4 * H)).astype(np.float32) for d in ("fw", "bw")}
    bias = {d: rng.normal(0, 0.1, size=4 * H).astype(np.float32) for d in ("fw", "bw")}
    pre = np.zeros((B * T, 8 * H), np.float32)
    for i, d in enumerate(("fw", "bw")):
        pre[:, i * 4 * H:(i + 1) * 4 * H] = x.reshape(B * T, D) @ kern[d][:D] + bias[d]
    out = np.full((B * T, 2 * H), -7.0, np.float32)
    whh = {d: np.ascontiguousarray(kern[d][D:]) for d in kern}
    rc = emu.emu_lstm(rg, B, T, H, _fp(pre), 8 * H, _fp(whh["fw"]), _fp(whh["bw"]), lens.ctypes.data_as(ctypes.c_void_p),
                      _fp(out), 2 * H)
    assert rc == 0
    ref = np.concatenate([O.lstm_direction(x, lens, kern[d], bias[d], d == "bw", np.float64) for d in ("fw", "bw")], axis=2)
    got = out.reshape(B, T, 2 * H)
    assert np.abs(got - ref).max() < 2e-5
    for b in range(B):
        assert (got[b, lens[b]:] == 0).all()


def test_head_path_prob_seq_len_and_greedy_kernels(emu, dna_model):
    """cb_head_decode_kernels.cuh: both layouts of the logit head, path_prob (warp shuffle reduction), seq_len scaling
    (round half to even in f64) and the greedy CTC decoder (shuffle / ballot / popc stream compaction, exact ties: first
    maximum wins) against the oracle."""
    cfg, t, _ = dna_model
    H, C = cfg.hidden, cfg.n_class
    rng = np.random.default_rng(12)
    B, T = 37, 45
    lasth = rng.normal(size=(B, T, 2 * H)).astype(np.float32)
    ref = O.head_forward(lasth, cfg, t)
    w = [np.ascontiguousarray(t["rnn_fnn_layer/" + n], dtype=np.float32) for n in ("weights", "bias", "weights_class", "bias_class")]
    got = np.zeros((B, T, C), np.float32)
    assert emu.emu_head(_fp(lasth), ctypes.c_longlong(B * T), H, C, _fp(w[0]), _fp(w[1]), _fp(w[2]), _fp(w[3]), _fp(got), 3) == 0
    assert np.abs(got - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    # time-major layout of the tensor-core LSTM stack: [T][2 x H/4][Bp][4]
    Bp = 128
    tm = np.zeros((T, 2 * (H // 4), Bp, 4), np.float32)
    tm[:, :, :B, :] = lasth.reshape(B, T, 2 * (H // 4), 4).transpose(1, 2, 0, 3)
    got_tm = np.zeros((B, T, C), np.float32)
    assert emu.emu_head_tmajor(_fp(tm), B, Bp, T, H, C, _fp(w[0]), _fp(w[1]), _fp(w[2]), _fp(w[3]), _fp(got_tm)) == 0
    assert np.abs(got_tm - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    # decoder inputs with exact ties
    lg = rng.normal(size=(B, T, C)).astype(np.float32)
    lg[:, :, C - 1] += 1.5
    lg[:, ::5, :] = np.round(lg[:, ::5, :])
    prob = np.zeros(B, np.float32)
    assert emu.emu_path_prob(_fp(lg), B, T, C, _fp(prob)) == 0
    np.testing.assert_allclose(prob, O.path_prob(lg), rtol=0, atol=2e-6)
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0], lens[1] = T, 0
    bases = np.full((B, T), 9, np.int8)
    n_bases = np.zeros(B, np.int32)
    assert emu.emu_greedy(_fp(lg), lens.ctypes.data_as(ctypes.c_void_p), B, T, C, bases.ctypes.data_as(ctypes.c_void_p),
                          n_bases.ctypes.data_as(ctypes.c_void_p)) == 0
    assert [bases[b, :n_bases[b]].tolist() for b in range(B)] == O.ctc_greedy(lg, lens)
    assert all((bases[b, n_bases[b]:] == 0).all() for b in range(B))
    for L, Tq in ((400, 400), (500, 100), (503, 101), (2000, 286)):
        raw = np.concatenate([np.arange(0, L + 1, 7), [L, 1, L - 1, 5 * (L // 10)]]).astype(np.int32)
        out = np.zeros_like(raw)
        assert emu.emu_seq_len(raw.ctypes.data_as(ctypes.c_void_p), len(raw), L, Tq, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.array_equal(out, O.seq_len_out(raw, L / Tq))


@pytest.mark.parametrize("warp", [1, 2, 0], ids=["beam_warp_kernel", "beam_warp_kernel_staged", "beam_kernel"])
def test_beam_search_kernels_are_bit_identical_to_the_c_oracle(emu, warp):
    """The warp-cooperative shared-memory beam search (the product path) and the thread-per-window fallback against the C
    oracle's restatement of TF's CTCBeamSearchDecoder: widths 1..30, ragged lengths incl. 0, exact ties, with the launcher's
    small pool (the trie is compacted in place when it fills up) and with the pool that can never overflow."""
    rng = np.random.default_rng(31)
    B, T, C = 9, 36, 5
    lg = rng.normal(size=(B, T, C)).astype(np.float32) * 2
    lg[:, :, C - 1] += 1.0
    lg[:, ::4, :] = np.round(lg[:, ::4, :])
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0], lens[1] = T, 0
    overflowed, compared = [], 0
    for W in (1, 2, 5, 13, 30):
        for pool in sorted({max(6 * W, 64), 2 * W * (T + 1) + 2}):      # cb_launch_beam's small pool, and the no-overflow bound
            bases = np.full((B, T), 9, np.int8)
            n_bases = np.full(B, -1, np.int32)
            rc = emu.emu_beam(warp, _fp(lg), lens.ctypes.data_as(ctypes.c_void_p), B, T, C, W, pool,
                              bases.ctypes.data_as(ctypes.c_void_p), n_bases.ctypes.data_as(ctypes.c_void_p))
            if pool < 2 * W * (T + 1) + 2 and rc == 1:
                overflowed.append(W)          # legitimate on random logits: cb_launch_beam then falls back / retries
                continue
            assert rc == 0, "error %d at W=%d pool=%d" % (rc, W, pool)
            compared += 1
            assert [bases[b, :n_bases[b]].tolist() for b in range(B)] == O.ctc_decode_c(lg, lens, W), (W, pool)
            assert all((bases[b, n_bases[b]:] == 0).all() for b in range(B))
    assert compared >= 7 and set(overflowed) <= {13, 30}      # random logits: wide beams may outgrow the small pool


@pytest.mark.parametrize("warp", [1, 2, 0], ids=["beam_warp_kernel", "beam_warp_kernel_staged", "beam_kernel"])
def test_beam_search_kernels_report_the_top_path_score(emu, warp):
    """cb_decode_beam_scored: the `log_prob` output of the serving signature (chiron/export_test.py:36-40,103-113) = the
    score TopPaths() reports for the decoded path.  Every kernel variant against the C oracle's restatement, bit for bit
    (same libm on the host), and against an independent check: at width 1 with no pruning pressure the score of an
    all-blank path is the sum of the per-frame max-subtracted blank logits."""
    rng = np.random.default_rng(77)
    B, T, C = 7, 30, 5
    lg = rng.normal(size=(B, T, C)).astype(np.float32) * 2
    lg[:, :, C - 1] += 1.5
    lg[6] = -8.0
    lg[6, :, C - 1] = 0.25                        # blank wins every frame by a wide margin
    lens = np.array([T, 0, 1, 17, T, 9, T], np.int32)
    vp = ctypes.c_void_p
    for W in (1, 4, 30):
        pool = 2 * W * (T + 1) + 2
        bases = np.full((B, T), 9, np.int8)
        n_bases = np.full(B, -1, np.int32)
        scores = np.full(B, np.nan, np.float32)
        assert emu.emu_beam_scored(warp, _fp(lg), lens.ctypes.data_as(vp), B, T, C, W, pool, bases.ctypes.data_as(vp),
                                   n_bases.ctypes.data_as(vp), _fp(scores)) == 0
        paths, ref = O.ctc_beam_scores_c(lg, lens, W)
        assert [bases[b, :n_bases[b]].tolist() for b in range(B)] == paths
        assert np.array_equal(scores, ref), (W, scores, ref)
        assert scores[1] == 0.0 and n_bases[6] == 0           # empty window: log(1); all-blank window decodes to nothing
    # all-blank path, width 1: only the blank extension survives, its score is the sum of (blank - max) = 0 per frame
    one = np.zeros(B, np.float32)
    emu.emu_beam_scored(warp, _fp(lg), lens.ctypes.data_as(vp), B, T, C, 1, 2 * (T + 1) + 2, bases.ctypes.data_as(vp),
                        n_bases.ctypes.data_as(vp), _fp(one))
    assert one[6] == 0.0


def test_assembly_kernels_reproduce_the_reference_fixtures(emu):
    """asm_compact -> asm_disp -> asm_scan -> asm_vote -> asm_finish (with the launcher's grids and workspace plan) on the
    fixtures produced by the reference's own easy_assembler + qs() (tests/golden/assembly_ref): consensus, quality string
    and window positions for the simple / glue / stick kernels, with empty windows mixed in."""
    from test_assembly_reference_fixtures import BASE_IDX, FIXTURES, _load
    code = {"simple": 0, "glue": 1, "stick": 2}
    for path in FIXTURES[:3]:
        fx, segs, wts = _load(path)
        T = max(len(s) for s in segs) + 3
        n = len(segs) + 2
        rows = [i for i in range(n) if i not in (4, n - 1)]         # two empty windows: dropped like sparse2dense does
        bases = np.zeros((n, T), np.int8)
        n_bases = np.zeros(n, np.int32)
        prob = np.zeros(n, np.float32)
        for r, sgm, q in zip(rows, segs, wts):
            bases[r, :len(sgm)] = [BASE_IDX[c] for c in sgm]
            n_bases[r] = len(sgm)
            prob[r] = q
        for case in fx["cases"]:
            L = 400
            jump = int(round(case["jump_step_ratio"] * L))
            max_len = int(n_bases.sum()) + 1
            cons = np.zeros(max_len, np.int8)
            qual = np.zeros(max_len, np.uint8)
            pos = np.full(n, -5, np.int32)
            out_len = np.zeros(1, np.int32)
            rc = emu.emu_assemble(bases.ctypes.data_as(ctypes.c_void_p), n_bases.ctypes.data_as(ctypes.c_void_p), _fp(prob), n, T,
                                  jump, L, code[case["kernal"]], cons.ctypes.data_as(ctypes.c_void_p),
                                  qual.ctypes.data_as(ctypes.c_void_p), pos.ctypes.data_as(ctypes.c_void_p),
                                  out_len.ctypes.data_as(ctypes.c_void_p), max_len)
            assert rc == 0
            ln = int(out_len[0])
            assert O.index2base(cons[:ln]) == case["consensus"], (os.path.basename(path), case["kernal"])
            ref_cons, _, ref_pos = O.simple_assembly_qs(segs, wts, case["jump_step_ratio"], kernal=case["kernal"])
            assert pos[rows].tolist() == ref_pos.tolist() and pos[4] == -1 and pos[n - 1] == -1
            covered = ref_cons.sum(axis=0) > 0
            got_q = bytes(qual[:ln]).decode("latin-1")
            assert [c for c, ok in zip(got_q, covered) if ok] == [c for c, ok in zip(case["quality"], covered) if ok]


def test_kernels_are_race_free_under_thread_sanitizer():
    """tests/cuda_emu/race_check.cpp: every SIMT kernel of the library (fp32 conv stack in both BatchNorm modes with and
    without stem, LSTM / GRU recurrences, head, path_prob, seq_len, greedy, the three beam-search kernels, the five assembly
    kernels) run under the host emulation with -fsanitize=thread.  The emulation's only synchronisation is what the
    kernel asks for (__syncthreads, __syncwarp, warp collectives, atomics), so a missing barrier -- what compute-sanitizer's
    racecheck looks for -- is a ThreadSanitizer data race.  A deliberately racy kernel proves the check can fail."""
    exe = os.path.join(HERE, "cuda_emu", "_build", "race_check")
    src = os.path.join(HERE, "cuda_emu", "race_check.cpp")
    deps = [src, SRC, os.path.join(HERE, "cuda_emu", "cuda_emu.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                                                         if f.endswith((".cuh", ".h"))]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-pthread", "-Wno-unknown-pragmas",
                               "-o", exe, src])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66")
    control = subprocess.run([exe, "--self-test"], capture_output=True, text=True, env=env, timeout=300)
    assert control.returncode == 66 and "ThreadSanitizer: data race" in control.stderr, "the race check cannot detect a race"
    run = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=900)
    assert "ThreadSanitizer" not in run.stderr, run.stderr[:4000]
    assert run.returncode == 0 and "race_check:" in run.stdout, (run.returncode, run.stdout[-500:], run.stderr[-2000:])


def test_kernels_are_memory_clean_under_address_sanitizer():
    """The same driver built with -fsanitize=address,undefined: global buffers are heap vectors and __shared__ arrays are
    statics with redzones, so an out-of-bounds or misaligned access of a kernel (compute-sanitizer memcheck's findings)
    aborts the run."""
    exe = os.path.join(HERE, "cuda_emu", "_build", "mem_check")
    src = os.path.join(HERE, "cuda_emu", "race_check.cpp")
    deps = [src, SRC, os.path.join(HERE, "cuda_emu", "cuda_emu.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                                                         if f.endswith((".cuh", ".h"))]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                               "-pthread", "-Wno-unknown-pragmas", "-o", exe, src])
    run = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and "race_check:" in run.stdout, (run.returncode, run.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[:3000]


def test_golden_windows_through_the_emulated_kernels(emu, dna_model):
    """The reference's golden output reproduced on the CPU by the CUDA kernel SOURCES: two windows of the bundled read1
    signal go through the emulated fp32 kernels end to end -- conv stack (folded population BN), hoisted input projections,
    the three BiLSTM layers, the logit head, beam search of width 30 -- and must decode to the reference's own
    segments/read1.fastq records (example_data/DNA/output), with logits within the fp32 tolerance of the oracle."""
    from conftest import GOLDEN, read_fasta_records
    cfg, t, _ = dna_model
    C, H = cfg.channels, cfg.hidden
    sig = O.read_signal_text(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), 400, 390)
    pick = [0, 160]                                   # the first window and the ragged last one of the read
    x, lens = np.ascontiguousarray(x[pick]), lens[pick]
    B, L = x.shape
    # conv stack with population BN folded exactly as cb_create does
    ptrs, rank1 = [], []
    for b in range(cfg.n_blocks):
        p = "res_layer%d" % (b + 1)
        has1 = cfg.branch1_bn_mask >> b & 1
        inv1, sh1 = _fold(t, p + "/branch1/conv1") if has1 else (np.ones(C, np.float32), np.zeros(C, np.float32))
        inva, sha = _fold(t, p + "/branch2/conv2a")
        invb, shb = _fold(t, p + "/branch2/conv2b")
        invc, shc = _fold(t, p + "/branch2/conv2c")
        w1, w2a = t[p + "/branch1/conv1/weights"], t[p + "/branch2/conv2a/weights"]
        w2b, w2c = t[p + "/branch2/conv2b/weights"], t[p + "/branch2/conv2c/weights"]
        if b == 0:
            rank1 = [w2a.reshape(-1), inva, sha, w1.reshape(-1), inv1, sh1]
            ptrs += [np.zeros(4, np.float32), np.zeros(4, np.float32), w2b.reshape(-1, C) * invb, shb, w2c * invc, shc]
        else:
            ptrs += [w2a * inva, sha, w2b.reshape(-1, C) * invb, shb, np.concatenate([w2c * invc, w1 * inv1]), shc + sh1]
    fea, _ = _run_stack(emu, 0, cfg, ptrs, rank1, x, 4)
    T = fea.shape[1]
    Z = np.ascontiguousarray(fea.reshape(B * T, C))
    vp = ctypes.c_void_p
    lens32 = lens.astype(np.int32)
    for l in range(cfg.n_layers):                     # stacked-bidirectional layout: one projection for both directions
        D = Z.shape[1]
        kern = {d: t["lstm/%d/%s/kernel" % (l, d)] for d in ("fw", "bw")}
        wx = np.ascontiguousarray(np.concatenate([kern["fw"][:D], kern["bw"][:D]], axis=1), dtype=np.float32)
        bias = np.concatenate([t["lstm/%d/fw/bias" % l], t["lstm/%d/bw/bias" % l]]).astype(np.float32)
        pre = np.zeros((B * T, 8 * H), np.float32)
        assert emu.emu_gemm(_fp(Z), D, _fp(wx), _fp(bias), B * T, 8 * H, D, 0, _fp(pre), 8 * H) == 0
        whh = {d: np.ascontiguousarray(kern[d][D:], dtype=np.float32) for d in kern}
        out = np.zeros((B * T, 2 * H), np.float32)
        assert emu.emu_lstm(1, B, T, H, _fp(pre), 8 * H, _fp(whh["fw"]), _fp(whh["bw"]), lens32.ctypes.data_as(vp), _fp(out),
                            2 * H) == 0
        Z = out
    hw = [np.ascontiguousarray(t["rnn_fnn_layer/" + n], dtype=np.float32) for n in ("weights", "bias", "weights_class", "bias_class")]
    logits = np.zeros((B, T, cfg.n_class), np.float32)
    assert emu.emu_head(_fp(Z), ctypes.c_longlong(B * T), H, cfg.n_class, _fp(hw[0]), _fp(hw[1]), _fp(hw[2]), _fp(hw[3]),
                        _fp(logits), 4) == 0
    ref = O.inference(x, lens, cfg, t)
    assert np.abs(logits - ref).max() < 2e-3
    W = 30
    bases = np.zeros((B, T), np.int8)
    n_bases = np.zeros(B, np.int32)
    assert emu.emu_beam(1, _fp(logits), lens32.ctypes.data_as(vp), B, T, cfg.n_class, W, 2 * W * (T + 1) + 2,
                        bases.ctypes.data_as(vp), n_bases.ctypes.data_as(vp)) == 0
    golden = read_fasta_records(os.path.join(GOLDEN, "DNA", "segments", "read1.fastq"))
    assert len(golden) == 161
    assert [O.index2base(bases[i, :n_bases[i]]) for i in range(B)] == [golden[k] for k in pick]


def test_decoders_on_the_reference_logits_sample(emu):
    """chiron/utils/logits_sample.npy (the reference's sample of real CTC logits; 24 of its 1100 windows are kept under
    tests/golden/logits): greedy and beam-search kernels (width 30 and 50, as the presets and the README use) under emulation
    against the C oracle, and the C oracle against the Python restatement of TF's decoder on a few windows."""
    lg = np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "logits", "logits_sample_24.npy"))
    B, T, C = lg.shape
    assert (B, T, C) == (24, 300, 5) and lg.dtype == np.float32
    rng = np.random.default_rng(2)
    lens = np.full(B, T, np.int32)
    lens[::5] = rng.integers(1, T, size=len(lens[::5]))
    vp = ctypes.c_void_p
    bases = np.zeros((B, T), np.int8)
    n_bases = np.zeros(B, np.int32)
    assert emu.emu_greedy(_fp(lg), lens.ctypes.data_as(vp), B, T, C, bases.ctypes.data_as(vp), n_bases.ctypes.data_as(vp)) == 0
    greedy = [bases[b, :n_bases[b]].tolist() for b in range(B)]
    assert greedy == O.ctc_greedy(lg, lens) == O.ctc_decode_c(lg, lens, 0)
    assert 10 < np.mean([len(g) for g in greedy]) < 40              # ~20 bases per 300-sample window on real data
    from chiron_b200 import _lib
    lib = _lib.load()
    out = np.zeros(T, np.int8)
    sub = [0, 5, 9, 14, 20, 23]                        # emulation is slow: six windows go through the kernels ...
    for W in (30, 50):
        ref = O.ctc_decode_c(lg, lens, W)
        pool = emu.emu_beam_small_pool(T, W)           # the first pass's pool: no window of these real logits overflows it
        # 16W nodes grown into the slack of the occupancy step: 4 CTAs (16 warps) per SM at W=30, 2 at W=50 (26 B per node)
        assert pool == (492 if W == 30 else 1014) and pool >= 16 * W
        small = 0
        for b in range(B):                             # ... and all 24 through the same search, host-compiled, with 32-bit
            row = np.ascontiguousarray(lg[b])          # (beam_kernel) and 16-bit (shared-memory kernels) trie nodes
            for fn in (lib.cb_selftest_beam, lib.cb_selftest_beam16):
                n = fn(row.ctypes.data_as(vp), int(lens[b]), C, W, pool, out.ctypes.data_as(vp))
                assert n >= 0 and out[:n].tolist() == ref[b], (W, b, n)
            small += lib.cb_selftest_beam16(row.ctypes.data_as(vp), int(lens[b]), C, W, 6 * W, out.ctypes.data_as(vp)) == -2
        assert small >= 3                              # why the pool is not 6W: real logits overflow it regularly
        lg_s, lens_s = np.ascontiguousarray(lg[sub]), np.ascontiguousarray(lens[sub])
        for warp in ((1, 2) if W == 30 else (1,)):
            bases[:] = 9
            assert emu.emu_beam(warp, _fp(lg_s), lens_s.ctypes.data_as(vp), len(sub), T, C, W, pool, bases.ctypes.data_as(vp),
                                n_bases.ctypes.data_as(vp)) == 0
            assert [bases[i, :n_bases[i]].tolist() for i in range(len(sub))] == [ref[b] for b in sub], (W, warp)
    for b in (0, 7, 23):
        assert O.ctc_beam_search_one(lg[b], int(lens[b]), 30) == O.ctc_decode_c(lg[b:b + 1], lens[b:b + 1], 30)[0]
    # wide beams get what four windows can hold in the shared-memory budget, never less than the algorithm's minimum
    assert 2 * 100 + 2 <= emu.emu_beam_small_pool(150, 100) < 24 * 100
    assert emu.emu_beam_small_pool(3, 30) == 2 * 30 * 4 + 2 and 64 <= emu.emu_beam_small_pool(300, 1) <= 128


def test_three_pass_beam_search_on_real_logits(emu):
    """cb_launch_beam's three passes exactly as it enqueues them, on the reference's logits sample: a first pass whose pool
    (here 6W nodes, so that windows DO overflow) marks the windows that outgrow it, beam_retry_kernel redoes those alone with
    a CTA-sized pool, beam_kernel's slot mode finishes what is still marked; together bit-identical to the C oracle.  With a
    starved retry pool the third pass does the work; with too few workspaces it raises the sticky error flag."""
    lg = np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "logits", "logits_sample_24.npy"))[:12]
    B, T, C = lg.shape
    lens = np.full(B, T, np.int32)
    lens[3] = 120
    W = 30
    vp = ctypes.c_void_p
    ref = O.ctc_decode_c(lg, lens, W)

    def run(pool_first, pool_retry, n_slots):
        bases = np.full((B, T), 9, np.int8)
        n_bases = np.zeros(B, np.int32)
        marked = (ctypes.c_int * 3)()
        rc = emu.emu_beam_passes(_fp(lg), lens.ctypes.data_as(vp), B, T, C, W, pool_first, pool_retry, n_slots,
                                 bases.ctypes.data_as(vp), n_bases.ctypes.data_as(vp), marked)
        return rc, list(marked), bases, n_bases

    rc, marked, bases, n_bases = run(6 * W, 0, 4)
    assert rc == 0 and 1 <= marked[0] < B and marked[1] == 0 and marked[2] == 0, (rc, marked)   # some overflow 6W, none the retry pool
    assert [bases[b, :n_bases[b]].tolist() for b in range(B)] == ref
    assert all((bases[b, n_bases[b]:] == 0).all() for b in range(B))
    rc, marked2, bases, n_bases = run(6 * W, 6 * W, B)             # a retry pool no larger than the first: the third pass decodes
    assert rc == 0 and marked2[0] == marked2[1] == marked[0] and marked2[2] == 0, (rc, marked2)
    assert [bases[b, :n_bases[b]].tolist() for b in range(B)] == ref
    if marked[0] >= 2:                                           # fewer workspaces than marked windows: sticky error, rest decoded
        rc, marked3, bases, n_bases = run(6 * W, 6 * W, 1)
        assert rc == 1 and marked3[2] == 0
        ok = [bases[b, :n_bases[b]].tolist() == ref[b] for b in range(B)]
        assert sum(ok) == B - (marked[0] - 1) and all(n_bases[b] == 0 for b in range(B) if not ok[b])
