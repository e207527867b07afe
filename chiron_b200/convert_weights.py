"""Offline converter: TF checkpoint folder (``chiron/model/<name>/``) -> packed CBW1 blob.

Replaces ``tf.train.Saver.restore(latest_checkpoint(model_dir))`` (chiron/chiron_eval.py:272-276).  The topology
(number of residual blocks, conv2b width and stride per block, RNN layout) is derived from the checkpoint tensor
shapes and the ``.meta`` graph rather than from ``model.json``, which is wrong for RNA_default (SURVEY.md finding 3).

    python -m chiron_b200.convert_weights /root/reference/chiron/model/DNA_default chiron_b200/weights/DNA_default.cbw
"""
from __future__ import annotations

import json
import os
import sys
from typing import Dict

import numpy as np

from . import tf_bundle
from .model import (BN_BATCH, BN_POPULATION, CELL_GRU, CELL_LSTM, ModelConfig, NORM_FULL_MAD, NORM_UNIQUE_MAD, RNN_NORMAL, RNN_RNA,
                    pack_blob)


def _take_bn(raw, tensors, bn_modes, p: str, conv: str, required: bool) -> bool:
    """Collect the BN variables of ``<p>/<conv>``.  Two variable sets exist.  The shipped checkpoints: batchnorm()'s
    <conv>_bn/{scale,offset,pop_mean,pop_var} (chiron/cnn.py:140-148) -> population mode.  A model trained at HEAD:
    simple_global_bn's <conv>_bn/<conv>_bn_{scale,offset} (chiron/cnn.py:65-68,181-186) and no statistics -> batch mode.
    Returns whether the convolution has BN at all."""
    leaf = conv.rsplit("/", 1)[-1]
    head_fmt = "%s/%s_bn/%s_bn_%%s" % (p, conv, leaf)
    has_pop = "%s/%s_bn/scale" % (p, conv) in raw
    has_head = head_fmt % "scale" in raw
    if required and not (has_pop or has_head):
        raise ValueError("%s/%s has no BN variables in the checkpoint" % (p, conv))
    if has_pop:
        bn_modes.add(BN_POPULATION)
        for n in ("scale", "offset", "pop_mean", "pop_var"):
            tensors["%s/%s_bn/%s" % (p, conv, n)] = raw["%s/%s_bn/%s" % (p, conv, n)]
    elif has_head:
        bn_modes.add(BN_BATCH)
        scale = np.asarray(raw[head_fmt % "scale"], dtype=np.float32).reshape(-1)
        tensors["%s/%s_bn/scale" % (p, conv)] = scale
        tensors["%s/%s_bn/offset" % (p, conv)] = np.asarray(raw[head_fmt % "offset"], np.float32).reshape(-1)
        tensors["%s/%s_bn/pop_mean" % (p, conv)] = np.zeros_like(scale)     # unused in batch mode
        tensors["%s/%s_bn/pop_var" % (p, conv)] = np.ones_like(scale)
    return has_pop or has_head


# CNN front ends of chiron/cnn.py:350-362 this package runs ("model" of model.json's "cnn" section); anything else is refused
# by name, with the reason, before the variables are even looked at.
SUPPORTED_CNN = {"dna_model1", "rna_model2", "rna_model3", "rna_test"}
UNRUNNABLE_CNN = {
    "rna_model1": "chiron/cnn.py:391-393 calls tf.nn.avg_pool(net, ksize, strides) without its required `padding` argument: "
                  "the graph cannot be built at HEAD, so no checkpoint of this topology can exist",
}


def check_cnn_name(model_json: dict):
    name = ((model_json or {}).get("cnn") or {}).get("model")
    if name is None or name in SUPPORTED_CNN:
        return
    if name in UNRUNNABLE_CNN:
        raise ValueError("CNN model %r is unrunnable in the reference itself: %s" % (name, UNRUNNABLE_CNN[name]))
    raise ValueError("CNN model %r (chiron/cnn.py:350-362) is a research topology without shipped weights and is out of "
                     "scope; supported: %s" % (name, ", ".join(sorted(SUPPORTED_CNN))))


def convert_tensors(raw: Dict[str, np.ndarray], conv_attrs: Dict[str, dict], model_json: dict) -> bytes:
    check_cnn_name(model_json)
    n_blocks = 0
    while "res_layer%d/branch2/conv2b/weights" % (n_blocks + 1) in raw:
        n_blocks += 1
    if n_blocks == 0:
        raise ValueError("checkpoint has no res_layerN/branch2/conv2b/weights")
    C = raw["res_layer1/branch2/conv2b/weights"].shape[-1]
    k, stride, mask = [], [], 0
    bn_modes = set()
    tensors: Dict[str, np.ndarray] = {}
    stem_k = stem_stride = 0
    if "conv_layer/conv1/weights" in raw:                      # RNA_model2 / RNA_model3 stem (chiron/cnn.py:454-476)
        w = raw["conv_layer/conv1/weights"]                    # (1, k, 1, C) HWIO
        stem_k = int(w.shape[1])
        attr = conv_attrs.get("conv_layer/conv1/conv1")
        if attr and attr["strides"]:
            stem_stride = int(attr["strides"][2])
        else:                                                  # no .meta: the two stems HEAD defines
            stem_stride = {9: 5, 14: 7}.get(stem_k, 0)
        if stem_stride < 1:
            raise ValueError("cannot determine the stride of conv_layer/conv1 (no .meta graph)")
        tensors["conv_layer/conv1/weights"] = w.reshape(stem_k, C)
    for b in range(n_blocks):
        p = "res_layer%d" % (b + 1)
        w2b = raw[p + "/branch2/conv2b/weights"]            # (1, k, C, C) HWIO
        k.append(int(w2b.shape[1]))
        s = 1
        attr = conv_attrs.get(p + "/branch2/conv2b/conv2b")
        if attr and attr["strides"]:
            s = int(attr["strides"][2])
            s1 = int(conv_attrs[p + "/branch1/conv1/conv1"]["strides"][2])
            if s1 != s:
                raise ValueError("block %d: branch1 stride %d != conv2b stride %d" % (b + 1, s1, s))
        stride.append(s)
        for conv in ("branch1/conv1", "branch2/conv2a", "branch2/conv2b", "branch2/conv2c"):
            w = raw["%s/%s/weights" % (p, conv)]
            tensors["%s/%s/weights" % (p, conv)] = w.reshape(w.shape[1:]) if conv.endswith("conv2b") \
                else w.reshape(w.shape[2:])
            has_bn = _take_bn(raw, tensors, bn_modes, p, conv, required=conv != "branch1/conv1")
            if conv == "branch1/conv1":
                mask |= int(has_bn) << b
    if stem_k:
        _take_bn(raw, tensors, bn_modes, "conv_layer", "conv1", required=True)
    if any(n.startswith("BDLSTM_rnn/") for n in raw):
        layout = RNN_NORMAL
        fmt = "BDLSTM_rnn/cell_{l}/bidirectional_rnn/{d}/lstm_cell/{t}"     # chiron/rnn.py:62-64
    elif any(n.startswith("BDGRU_rnn/") for n in raw):
        layout = RNN_RNA
        fmt = "BDGRU_rnn/{d}/multi_rnn_cell/cell_{l}/lstm_cell/{t}"        # chiron/rnn.py:140-143
    else:
        raise ValueError("checkpoint has no LSTM variables")
    # LSTMCell: .../lstm_cell/{kernel,bias}; GRUCell: .../gru_cell/{gates,candidate}/{kernel,bias} (chiron/rnn.py:47-53)
    cell_type = CELL_GRU if fmt.replace("lstm_cell", "gru_cell").format(l=0, d="fw", t="gates/kernel") in raw else CELL_LSTM
    if cell_type == CELL_GRU:
        fmt = fmt.replace("lstm_cell", "gru_cell")
    first = "gates/kernel" if cell_type == CELL_GRU else "kernel"
    n_layers = 0
    while fmt.format(l=n_layers, d="fw", t=first) in raw:
        n_layers += 1
    if n_layers == 0:
        raise ValueError("checkpoint has no LSTMCell / GRUCell variables (BNLSTM is not supported)")
    if cell_type == CELL_GRU:
        H = raw[fmt.format(l=0, d="fw", t="candidate/bias")].shape[0]
        for l in range(n_layers):
            for d in ("fw", "bw"):
                for leaf in ("gates/kernel", "gates/bias", "candidate/kernel", "candidate/bias"):
                    tensors["gru/%d/%s/%s" % (l, d, leaf)] = raw[fmt.format(l=l, d=d, t=leaf)]
    else:
        H = raw[fmt.format(l=0, d="fw", t="bias")].shape[0] // 4
        for l in range(n_layers):
            for d in ("fw", "bw"):
                tensors["lstm/%d/%s/kernel" % (l, d)] = raw[fmt.format(l=l, d=d, t="kernel")]
                tensors["lstm/%d/%s/bias" % (l, d)] = raw[fmt.format(l=l, d=d, t="bias")]
    for n in ("weights", "bias", "weights_class", "bias_class"):
        tensors["rnn_fnn_layer/" + n] = raw["rnn_fnn_layer/" + n]
    n_class = raw["rnn_fnn_layer/bias_class"].shape[0]
    rnn_json = model_json.get("rnn", {})
    if rnn_json.get("layer_num", n_layers) != n_layers or rnn_json.get("hidden_num", H) != H:
        raise ValueError("model.json rnn section disagrees with the checkpoint")
    if rnn_json.get("cell_type", "LSTM") not in ("LSTM", "GRU") or (rnn_json.get("cell_type", "LSTM") == "GRU") != (cell_type == CELL_GRU):
        raise ValueError("model.json cell_type %r disagrees with the checkpoint or is unsupported (LSTM and GRU cells only, "
                         "chiron/rnn.py:47-60)" % rnn_json.get("cell_type"))
    # Input normalisation is a property of how the weights were trained (SURVEY.md finding 4): DNA_default needs
    # the unique-value median/MAD (pinned by the golden outputs); RNA_default a scale~1 normalisation (unpinned).
    sig_norm = NORM_UNIQUE_MAD if layout == RNN_NORMAL else NORM_FULL_MAD
    if len(bn_modes) != 1:
        raise ValueError("checkpoint mixes population-statistics and batch-statistics BatchNorm variables")
    cfg = ModelConfig(n_blocks=n_blocks, channels=int(C), hidden=int(H), n_layers=n_layers, n_class=int(n_class),
                      rnn_layout=layout, branch1_bn_mask=mask, k=k, stride=stride, sig_norm=sig_norm,
                      reverse_signal=int(layout == RNN_RNA), bn_mode=bn_modes.pop(), cell_type=cell_type,
                      stem_k=stem_k, stem_stride=stem_stride)
    return pack_blob(cfg, tensors)


def convert_checkpoint_dir(model_dir: str) -> bytes:
    prefix = tf_bundle.latest_checkpoint(model_dir)
    raw = tf_bundle.read_checkpoint(prefix)
    attrs = tf_bundle.read_conv_attrs(prefix + ".meta") if os.path.exists(prefix + ".meta") else {}
    mj = {}
    if os.path.exists(os.path.join(model_dir, "model.json")):
        with open(os.path.join(model_dir, "model.json")) as f:
            mj = json.load(f)
    return convert_tensors(raw, attrs, mj)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print(__doc__)
        return 2
    blob = convert_checkpoint_dir(argv[0])
    os.makedirs(os.path.dirname(os.path.abspath(argv[1])), exist_ok=True)
    with open(argv[1], "wb") as f:
        f.write(blob)
    print("wrote %s (%d bytes)" % (argv[1], len(blob)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
