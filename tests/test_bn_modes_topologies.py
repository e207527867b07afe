"""CPU tests of SURVEY.md 8f-3 / 8f-4: the batch-statistics BatchNorm mode (HEAD's simple_global_bn,
chiron/cnn.py:166-188) and residual stacks other than the two shipped ones (e.g. the 5-block rna_test,
chiron/cnn.py:555-566).

No checkpoint trained at HEAD or with those topologies ships with the reference, so the oracle's restatement of them is
"parity unpinned" against the reference itself; what pins it here is an INDEPENDENT restatement with torch's own
conv1d / batch_norm / LSTM-free primitives on CPU (torch is not used by the oracle), plus structural identities."""
import os

import numpy as np
import pytest

from chiron_b200 import model as M
from chiron_b200.convert_weights import convert_tensors
from oracle import chiron_oracle as O


def _torch_cnn(x, cfg, t, bn_mode):
    """The residual stack written with torch.nn.functional (float64): F.conv1d on explicitly 'SAME'-padded input and
    F.batch_norm with training=True (batch moments, biased variance = tf.nn.moments) or the stored statistics."""
    import torch
    import torch.nn.functional as F

    def conv(inp, w, stride):                     # inp [B,Cin,T], w [k,Cin,Cout] (TF HWIO without the H axis)
        k = w.shape[0]
        T = inp.shape[2]
        t_out = -(-T // stride)
        pad = max((t_out - 1) * stride + k - T, 0)
        inp = F.pad(inp, (pad // 2, pad - pad // 2))
        return F.conv1d(inp, torch.from_numpy(np.array(w.transpose(2, 1, 0), dtype=np.float64)), stride=stride)

    def bn(inp, prefix):
        g = lambda n: torch.from_numpy(np.array(t[prefix + "_bn/" + n], dtype=np.float64))
        if bn_mode == 1:
            return F.batch_norm(inp, None, None, g("scale"), g("offset"), training=True, eps=1e-5)
        return F.batch_norm(inp, g("pop_mean"), g("pop_var"), g("scale"), g("offset"), training=False, eps=1e-5)

    net = torch.from_numpy(np.asarray(x, dtype=np.float64))[:, None, :]
    if cfg.stem_k:
        net = torch.relu(bn(conv(net, t["conv_layer/conv1/weights"][:, None, :], cfg.stem_stride), "conv_layer/conv1"))
    for b in range(cfg.n_blocks):
        p = "res_layer%d" % (b + 1)
        s = cfg.stride[b]
        b1 = conv(net, t[p + "/branch1/conv1/weights"][None], s)
        if cfg.branch1_bn_mask >> b & 1:
            b1 = bn(b1, p + "/branch1/conv1")
        a = torch.relu(bn(conv(net, t[p + "/branch2/conv2a/weights"][None], 1), p + "/branch2/conv2a"))
        bb = torch.relu(bn(conv(a, t[p + "/branch2/conv2b/weights"], s), p + "/branch2/conv2b"))
        c = bn(conv(bb, t[p + "/branch2/conv2c/weights"][None], 1), p + "/branch2/conv2c")
        net = torch.relu(b1 + c)
    return net.permute(0, 2, 1).numpy()


def _signal(B, L, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(-0.16, 0.43, size=(B, L)).astype(np.float32)
    x[1, L // 2:] = 0                              # a zero-padded short window: padding counts in the batch moments
    return x


@pytest.mark.parametrize("bn_mode", [0, 1])
def test_oracle_cnn_matches_torch_on_dna_default(dna_model, bn_mode):
    cfg, t, _ = dna_model
    x = _signal(5, 96)
    ref = _torch_cnn(x, cfg, t, bn_mode)
    got = O.cnn_forward(x, cfg, t, np.float64, bn_mode)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    got32 = O.cnn_forward(x, cfg, t, np.float32, bn_mode)
    assert np.abs(got32 - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("bn_mode", [0, 1])
def test_oracle_cnn_matches_torch_on_rna_default(rna_model, bn_mode):
    """k=13 / stride-5 first block: the strided 'SAME' padding and the strided 1x1 branch."""
    cfg, t, _ = rna_model
    x = _signal(3, 203, seed=1)                    # 203 is not a multiple of 5: asymmetric padding
    ref = _torch_cnn(x, cfg, t, bn_mode)
    got = O.cnn_forward(x, cfg, t, np.float64, bn_mode)
    assert got.shape == (3, 41, cfg.channels) == ref.shape
    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("k,stride,mask,stem", [([3] * 5, [1] * 5, 1, (0, 0)),            # rna_test (cnn.py:555-566)
                                                ([5, 3, 7, 3], [2, 1, 3, 1], 0b0101, (0, 0)),
                                                ([1, 2], [1, 2], 0b11, (0, 0)),
                                                ([3] * 3, [1] * 3, 1, (9, 5)),             # RNA_model2 (cnn.py:454-464)
                                                ([3] * 3, [1] * 3, 1, (14, 7))])           # RNA_model3 (cnn.py:466-476)
def test_oracle_cnn_matches_torch_on_other_topologies(k, stride, mask, stem):
    cfg = M.ModelConfig(n_blocks=len(k), channels=32, hidden=12, k=k, stride=stride, branch1_bn_mask=mask,
                        stem_k=stem[0], stem_stride=stem[1])
    t = M.random_tensors(cfg, seed=3)
    x = _signal(4, 77, seed=2)
    for bn_mode in (0, 1):
        ref = _torch_cnn(x, cfg, t, bn_mode)
        got = O.cnn_forward(x, cfg, t, np.float64, bn_mode)
        assert got.shape == (4, cfg.out_len(77), 32) == ref.shape
        assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_batch_mode_equals_population_mode_fed_with_the_batch_moments(dna_model):
    """Structural identity: storing the moments batch mode computed as pop_mean / pop_var and running population mode
    reproduces the batch-mode result exactly; and batch mode really differs from the shipped statistics."""
    cfg, t, _ = dna_model
    x = _signal(6, 120)
    stats = {}
    f_batch = O.cnn_forward(x, cfg, t, np.float32, 1, stats)
    assert len(stats) == 3 * cfg.n_blocks + bin(cfg.branch1_bn_mask).count("1")
    t2 = dict(t)
    for prefix, (mean, var) in stats.items():
        t2[prefix + "_bn/pop_mean"], t2[prefix + "_bn/pop_var"] = mean, var
    assert np.array_equal(O.cnn_forward(x, cfg, t2, np.float32, 0), f_batch)
    assert np.abs(O.cnn_forward(x, cfg, t, np.float32, 0) - f_batch).max() > 1e-2
    # a window's batch-mode result depends on its batch (HEAD behaviour) -- population mode's does not
    assert np.abs(O.cnn_forward(x[:3], cfg, t, np.float32, 1) - f_batch[:3]).max() > 1e-4
    assert np.array_equal(O.cnn_forward(x[:3], cfg, t, np.float32, 0), O.cnn_forward(x, cfg, t, np.float32, 0)[:3])


def test_blob_header_carries_bn_mode():
    cfg = M.ModelConfig(n_blocks=2, channels=8, hidden=4, k=[3, 5], stride=[1, 2], bn_mode=M.BN_BATCH)
    t = M.random_tensors(cfg, 0)
    cfg2, t2 = M.unpack_blob(M.pack_blob(cfg, t))
    assert cfg2 == cfg and cfg2.bn_mode == M.BN_BATCH
    assert all(np.array_equal(t[n], t2[n]) for n in t)
    assert M.unpack_blob(M.pack_blob(M.ModelConfig(), M.random_tensors(M.ModelConfig(), 0)))[0].bn_mode == M.BN_POPULATION


def _tf_checkpoint_names(cfg, t, head_literal):
    """The variable set tf.train.Saver would hold for this model: batchnorm()'s names (the shipped checkpoints,
    chiron/cnn.py:140-148) or simple_global_bn's (a model trained at HEAD, chiron/cnn.py:65-68,181-186); LSTM or GRU
    cells in either RNN layout; with or without the stem convolution."""
    raw = {}
    convs = [("res_layer%d" % (b + 1), conv) for b in range(cfg.n_blocks)
             for conv in ("branch1/conv1", "branch2/conv2a", "branch2/conv2b", "branch2/conv2c")]
    if cfg.stem_k:
        convs.insert(0, ("conv_layer", "conv1"))
    for p, conv in convs:
        w = t["%s/%s/weights" % (p, conv)]
        if p == "conv_layer":
            raw["%s/%s/weights" % (p, conv)] = w[None, :, None, :]               # (1, k, 1, C) HWIO
        else:
            raw["%s/%s/weights" % (p, conv)] = w[None] if conv.endswith("conv2b") else w[None, None]
        if "%s/%s_bn/scale" % (p, conv) not in t:
            continue
        leaf = conv.rsplit("/", 1)[-1]
        if head_literal:
            raw["%s/%s_bn/%s_bn_scale" % (p, conv, leaf)] = t["%s/%s_bn/scale" % (p, conv)]
            raw["%s/%s_bn/%s_bn_offset" % (p, conv, leaf)] = t["%s/%s_bn/offset" % (p, conv)]
        else:
            for n in ("scale", "offset", "pop_mean", "pop_var"):
                raw["%s/%s_bn/%s" % (p, conv, n)] = t["%s/%s_bn/%s" % (p, conv, n)]
    gru = cfg.cell_type == M.CELL_GRU
    cell = "gru_cell" if gru else "lstm_cell"
    fmt = ("BDLSTM_rnn/cell_{l}/bidirectional_rnn/{d}/%s/{t}" if cfg.rnn_layout == M.RNN_NORMAL
           else "BDGRU_rnn/{d}/multi_rnn_cell/cell_{l}/%s/{t}") % cell
    leaves = ("gates/kernel", "gates/bias", "candidate/kernel", "candidate/bias") if gru else ("kernel", "bias")
    for l in range(cfg.n_layers):
        for d in ("fw", "bw"):
            for leaf in leaves:
                raw[fmt.format(l=l, d=d, t=leaf)] = t["%s/%d/%s/%s" % ("gru" if gru else "lstm", l, d, leaf)]
    for n in ("weights", "bias", "weights_class", "bias_class"):
        raw["rnn_fnn_layer/" + n] = t["rnn_fnn_layer/" + n]
    return raw


def test_converter_maps_both_bn_variable_sets():
    cfg = M.ModelConfig(n_blocks=5, channels=16, hidden=8, k=[3] * 5, stride=[1] * 5, branch1_bn_mask=1)
    t = M.random_tensors(cfg, 5)
    c_pop, t_pop = M.unpack_blob(convert_tensors(_tf_checkpoint_names(cfg, t, False), {}, {}))
    assert c_pop == cfg and all(np.array_equal(t[n], t_pop[n]) for n in t)
    c_head, t_head = M.unpack_blob(convert_tensors(_tf_checkpoint_names(cfg, t, True), {}, {}))
    assert c_head.bn_mode == M.BN_BATCH and c_head.n_blocks == 5 and c_head.branch1_bn_mask == 1
    for n in t:
        if n.endswith("pop_mean"):
            assert not t_head[n].any()
        elif n.endswith("pop_var"):
            assert (t_head[n] == 1).all()
        else:
            assert np.array_equal(t[n], t_head[n])
    x = _signal(3, 40)
    assert np.array_equal(O.cnn_forward(x, c_head, t_head), O.cnn_forward(x, cfg, t, bn_mode=1))   # header's mode is used
    mixed = _tf_checkpoint_names(cfg, t, False)
    mixed.update({k: v for k, v in _tf_checkpoint_names(cfg, t, True).items() if "res_layer2" in k and "_bn_" in k})
    for k in [k for k in mixed if "res_layer2" in k and k.rsplit("/", 1)[-1] in ("scale", "offset", "pop_mean", "pop_var")]:
        del mixed[k]
    with pytest.raises(ValueError):
        convert_tensors(mixed, {}, {})


def _tf_gru_literal(x, lens, gk, gb, ck, cb, reverse):
    """rnn_cell_impl.GRUCell.call written literally with torch (concat, one matmul per gate set) inside dynamic_rnn's
    sequence_length rule -- no hoisting, no kernel splitting (the oracle does both)."""
    import torch
    B, T, D = x.shape
    H = cb.shape[0]
    xt = torch.from_numpy(np.asarray(x, np.float64))
    gk, gb, ck, cb = (torch.from_numpy(np.array(a, dtype=np.float64)) for a in (gk, gb, ck, cb))
    out = torch.zeros(B, T, H, dtype=torch.float64)
    for b in range(B):
        n = int(lens[b])
        seq = xt[b, :n].flip(0) if reverse else xt[b, :n]          # array_ops.reverse_sequence over the first len frames
        h = torch.zeros(H, dtype=torch.float64)
        hs = []
        for t in range(n):
            value = torch.sigmoid(torch.cat([seq[t], h]) @ gk + gb)
            r, u = value[:H], value[H:]
            c = torch.tanh(torch.cat([seq[t], r * h]) @ ck + cb)
            h = u * h + (1 - u) * c
            hs.append(h)
        if n:
            ys = torch.stack(hs)
            out[b, :n] = ys.flip(0) if reverse else ys
    return out.numpy()


@pytest.mark.parametrize("reverse", [False, True])
def test_oracle_gru_matches_literal_tf_formulation(reverse):
    rng = np.random.default_rng(9)
    B, T, D, H = 5, 11, 6, 8
    x = rng.normal(size=(B, T, D)).astype(np.float32)
    lens = np.array([T, 0, 1, 7, T - 1], dtype=np.int32)
    gk = rng.uniform(-0.5, 0.5, size=(D + H, 2 * H)).astype(np.float32)
    gb = (rng.normal(0, 0.1, size=2 * H) + 1).astype(np.float32)
    ck = rng.uniform(-0.5, 0.5, size=(D + H, H)).astype(np.float32)
    cb = rng.normal(0, 0.1, size=H).astype(np.float32)
    ref = _tf_gru_literal(x, lens, gk, gb, ck, cb, reverse)
    got = O.gru_direction(x, lens, gk, gb, ck, cb, reverse, np.float64)
    assert np.abs(got - ref).max() < 1e-12
    assert np.abs(O.gru_direction(x, lens, gk, gb, ck, cb, reverse, np.float32) - ref).max() < 1e-5


def test_gru_models_pack_convert_and_run_in_the_oracle():
    for layout in (M.RNN_NORMAL, M.RNN_RNA):
        cfg = M.ModelConfig(n_blocks=2, channels=16, hidden=8, n_layers=2, k=[3, 3], stride=[1, 1], rnn_layout=layout,
                            cell_type=M.CELL_GRU)
        t = M.random_tensors(cfg, 4)
        cfg2, t2 = M.unpack_blob(M.pack_blob(cfg, t))
        assert cfg2 == cfg and cfg2.cell_type == M.CELL_GRU and sorted(t2) == sorted(t)
        x = _signal(3, 30)
        lens = np.array([30, 12, 0], dtype=np.int32)
        lg = O.inference(x, lens, cfg, t)
        assert lg.shape == (3, 30, 5) and np.isfinite(lg).all()
        raw = _tf_checkpoint_names(cfg, t, False)          # TF variable names of a GRU checkpoint -> the same blob
        cfg3, t3 = M.unpack_blob(convert_tensors(raw, {}, {"rnn": {"cell_type": "GRU", "layer_num": 2, "hidden_num": 8}}))
        assert cfg3.cell_type == M.CELL_GRU and cfg3.rnn_layout == layout
        assert all(np.array_equal(t[n], t3[n]) for n in t)
        with pytest.raises(ValueError):
            convert_tensors(raw, {}, {"rnn": {"cell_type": "LSTM"}})


@pytest.mark.parametrize("head_literal", [False, True])
def test_converter_recognises_the_stem_convolution(head_literal):
    """RNA_model2 / RNA_model3 checkpoints (chiron/cnn.py:454-476): conv_layer/conv1 in front of the residual blocks."""
    cfg = M.ModelConfig(n_blocks=3, channels=16, hidden=8, k=[3] * 3, stride=[1] * 3, branch1_bn_mask=1, stem_k=14,
                        stem_stride=7)
    t = M.random_tensors(cfg, 6)
    assert t["res_layer1/branch1/conv1/weights"].shape == (16, 16)        # block 1 reads C channels behind a stem
    raw = _tf_checkpoint_names(cfg, t, head_literal)
    attrs = {"conv_layer/conv1/conv1": {"strides": [1, 1, 7, 1]}}
    for a in (attrs, {}):                                                 # with the .meta attrs, and from the kernel width alone
        cfg2, t2 = M.unpack_blob(convert_tensors(raw, a, {}))
        assert (cfg2.stem_k, cfg2.stem_stride, cfg2.n_blocks) == (14, 7, 3)
        assert cfg2.bn_mode == (M.BN_BATCH if head_literal else M.BN_POPULATION)
        assert np.array_equal(t2["conv_layer/conv1/weights"], t["conv_layer/conv1/weights"])
        assert cfg2.out_len(2000) == 286 and cfg2.total_stride() == 7
    x = _signal(3, 75)
    want = O.cnn_forward(x, cfg, t, bn_mode=int(head_literal))
    assert np.array_equal(O.cnn_forward(x, cfg2, t2), want) and want.shape == (3, 11, 16)


def test_rna_model1_is_unrunnable_in_the_reference_and_says_so():
    """`rna_model1` (chiron/cnn.py:391-401) starts with tf.nn.avg_pool(net, ksize, strides) WITHOUT the
    required `padding` argument, so the reference cannot even build that graph at HEAD and no checkpoint of it can exist.
    The converter refuses it by name and says why (it does not silently skip it); research topologies without shipped
    weights are refused as out of scope; the names the package runs pass."""
    import re
    from chiron_b200.convert_weights import SUPPORTED_CNN, check_cnn_name, convert_tensors
    with pytest.raises(ValueError, match="unrunnable in the reference.*avg_pool.*padding"):
        convert_tensors({}, {}, {"cnn": {"model": "rna_model1"}})
    for name in ("res_x", "variant_wavnet", "incp_v2", "custom", "gate_conv_net", "dynamic_net"):
        with pytest.raises(ValueError, match="out of scope"):
            check_cnn_name({"cnn": {"model": name}})
    for name in sorted(SUPPORTED_CNN) + [None]:
        check_cnn_name({"cnn": {"model": name}} if name else {})
    ref_cnn = os.path.join("/root/reference", "chiron", "cnn.py")
    if os.path.exists(ref_cnn):                       # in the build container: the claim, checked against the reference source
        src = open(ref_cnn).read()
        body = src[src.index("def RNA_model1"):src.index("def dynamic_net")]
        call = re.search(r"tf\.nn\.avg_pool\(([^)]*)\)", body).group(1)
        assert "padding" not in call
