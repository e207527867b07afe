"""GPU parity of GRU-cell models and stem-convolution models (model.json "cell_type": "GRU", chiron/rnn.py:51-53,129-131) through the
C ABI against the CPU oracle, on random-init weights -- no GRU checkpoint ships with the reference, so this is "parity
unpinned" against the reference; the oracle's GRUCell is pinned to a literal restatement of TF's cell in
test_bn_modes_topologies.py and the kernel source to the oracle under host emulation in test_cuda_emu.py."""
import os

import numpy as np
import pytest

from chiron_b200 import model as M
from oracle import chiron_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", [M.RNN_NORMAL, M.RNN_RNA])
def test_gru_model_matches_oracle_fp32(tmp_path, layout):
    from chiron_b200 import _lib
    from chiron_b200.engine import Basecaller
    cfg = M.ModelConfig(n_blocks=3, channels=256, hidden=100, n_layers=3, k=[3, 3, 3], stride=[1, 1, 1], branch1_bn_mask=1,
                        rnn_layout=layout, cell_type=M.CELL_GRU)
    t = M.random_tensors(cfg, seed=13)
    path = os.path.join(str(tmp_path), "gru.cbw")
    with open(path, "wb") as f:
        f.write(M.pack_blob(cfg, t))
    rng = np.random.default_rng(8)
    B, L = 70, 90                                   # two CTAs of 64 rows per direction, the second one partial
    x = rng.normal(-0.16, 0.43, size=(B, L)).astype(np.float32)
    lens = rng.integers(1, L + 1, size=B).astype(np.int32)
    lens[:3] = (L, 1, L - 1)
    for b in range(B):
        x[b, lens[b]:] = 0
    ref = O.inference(x, lens, cfg, t, np.float64)
    lasth = O.rnn_forward(O.cnn_forward(x, cfg, t, np.float64), lens, cfg, t, np.float64)
    bc = Basecaller(path, device=0, precision="fp32")
    bases, n_bases, prob, logits = bc.basecall_batch(x, lens, beam=0, want_logits=True)
    got_h = bc.debug_fetch(cfg.n_layers, lasth.size).reshape(lasth.shape)
    assert np.abs(got_h - lasth).max() < 1e-3
    for b in range(B):
        assert (got_h[b, lens[b]:] == 0).all()      # dynamic_rnn: zero output past sequence_length
    assert np.abs(logits - ref).max() < 2e-3
    s = np.sort(ref, axis=2)
    margin = s[:, :, -1] - s[:, :, -2]
    ref_paths = O.ctc_greedy(ref, lens)
    for b in range(B):
        if margin[b, :lens[b]].min() > 8e-3:
            assert bases[b, :n_bases[b]].tolist() == ref_paths[b]
    bc.close()
    with pytest.raises(_lib.ChironB200Error):       # the tensor-core recurrence is an LSTM kernel: refused loudly
        Basecaller(path, device=0, precision="tc")


@pytest.mark.parametrize("bn_mode", ["population", "batch"])
def test_stem_convolution_model_matches_oracle_fp32(tmp_path, bn_mode):
    """RNA_model3-shaped model (chiron/cnn.py:466-476): 1x14 stride-7 stem conv + BN + ReLU, then three residual blocks that
    all read 256 channels; random-init weights, fp32 path, both BatchNorm modes."""
    from chiron_b200 import _lib
    from chiron_b200.engine import Basecaller
    cfg = M.ModelConfig(n_blocks=3, channels=256, hidden=100, n_layers=3, k=[3, 3, 3], stride=[1, 1, 1], branch1_bn_mask=1,
                        rnn_layout=M.RNN_RNA, stem_k=14, stem_stride=7)
    t = M.random_tensors(cfg, seed=17)
    path = os.path.join(str(tmp_path), "rna_model3.cbw")
    with open(path, "wb") as f:
        f.write(M.pack_blob(cfg, t))
    rng = np.random.default_rng(9)
    B, L = 20, 500
    x = rng.normal(0.15, 0.9, size=(B, L)).astype(np.float32)
    lens = rng.integers(1, L + 1, size=B).astype(np.int32)
    lens[:2] = (L, 3)
    for b in range(B):
        x[b, lens[b]:] = 0
    T = cfg.out_len(L)
    assert T == 72
    lens_o = O.seq_len_out(lens, L / T)
    mode = M.BN_BATCH if bn_mode == "batch" else M.BN_POPULATION
    ref_fea = O.cnn_forward(x, cfg, t, np.float64, bn_mode=mode)
    ref = O.inference(x, lens_o, cfg, t, np.float64, bn_mode=mode)
    bc = Basecaller(path, device=0, precision="fp32", bn_mode=bn_mode)
    assert bc.out_len(L) == T and bc.out_len(2000) == 286
    bases, n_bases, prob, logits = bc.basecall_batch(x, lens, beam=0, want_logits=True)
    fea = bc.debug_fetch(0, ref_fea.size).reshape(ref_fea.shape)
    assert np.abs(fea - ref_fea).max() < 1e-3 * max(1.0, np.abs(ref_fea).max())
    assert np.abs(logits - ref).max() < (1e-2 if bn_mode == "batch" else 2e-3)
    bc.close()
    if bn_mode == "batch":
        with pytest.raises(_lib.ChironB200Error):   # batch statistics live on the fp32 kernels: refused loudly
            Basecaller(path, device=0, precision="tc", bn_mode="batch")
        return
    # the same blob on the tensor-core kernels: the stem convolution writes the operand image, and the first block is an
    # ordinary 256-channel block (cb_forward_tc); "auto" must pick them
    bt = Basecaller(path, device=0, precision="auto")
    assert bt.precision == "tc" and bt.out_len(L) == T
    bases_t, n_t, prob_t, logits_t = bt.basecall_batch(x, lens, beam=0, want_logits=True)
    fea_t = bt.debug_fetch(0, ref_fea.size).reshape(ref_fea.shape)
    assert np.abs(fea_t - ref_fea).max() < 2e-3 * max(1.0, np.abs(ref_fea).max())
    assert np.abs(logits_t - ref).max() < 5e-3
    assert np.array_equal(n_t, n_bases) and np.array_equal(bases_t, bases)       # greedy bases: tc == fp32 kernels
    bt.close()
