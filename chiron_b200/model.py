"""Model description + packed weight blob ("CBW1") shared by the oracle, the converter and the C-ABI.

The reference describes a model with ``model.json`` (chiron/chiron_model.py:37-48) plus a TF checkpoint restored by
``tf.train.Saver`` (chiron/chiron_eval.py:272-276).  ``model.json`` is not trustworthy for RNA_default (SURVEY.md
finding 3), so the topology here is *derived from checkpoint shapes and the .meta graph* by the converter and stored in
the blob header.  Blob layout (little endian):

    char[4]  "CBW1"
    int32    version (=1)
    int32    n_blocks, channels, hidden, n_layers, n_class, rnn_layout (0 = stacked-bidirectional "normal",
             1 = per-direction MultiRNNCell "rna"), branch1_bn_mask (bit b = block b's branch1 conv has BN)
    int32    k[8], stride[8]          (conv2b kernel width / stride of block b; branch1 shares the stride)
    int32    sig_norm (0 none, 1 unique-median/MAD, 2 full-signal median/MAD), reverse_signal,
             bn_mode (0 = population statistics, the shipped checkpoints' tf.cond BN, chiron/cnn.py:125-163;
             1 = batch statistics, HEAD's simple_global_bn, chiron/cnn.py:166-188),
             cell_type (0 = LSTMCell, 1 = GRUCell; chiron/rnn.py:47-53,126-131),
             stem_k, stem_stride (0, 0 = no stem; otherwise the strided 1 x stem_k convolution + BN + ReLU of the raw
             signal in front of the residual blocks: RNA_model2 = 9 / 5, RNA_model3 = 14 / 7, chiron/cnn.py:454-476)
    int64    n_floats
    float32  weights[n_floats]       in the canonical order of ``tensor_specs``

Canonical tensor order: with a stem, ``conv_layer/conv1`` W[stem_k,C] + bn first (and block 1 then reads C channels
instead of the one-channel signal); for every block ``branch1/conv1`` W[cin,C] (+bn), ``conv2a`` W[cin,C] +bn, ``conv2b``
W[k,C,C] +bn, ``conv2c`` W[C,C] +bn, where bn = scale, offset, pop_mean, pop_var (each [C]); then for every LSTM layer
and direction (fw, bw) kernel[in+H,4H] and bias[4H] (TF LSTMCell layout, gate column order i,j,f,o) -- or, for GRU cells,
gates/kernel[in+H,2H], gates/bias[2H] (columns r,u), candidate/kernel[in+H,H], candidate/bias[H] (TF GRUCell) --; then the head
``weights[2,H]``, ``bias[H]``, ``weights_class[H,n_class]``, ``bias_class[n_class]`` (chiron/rnn.py:73-88).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

MAGIC = b"CBW1"
MAX_BLOCKS = 8
BN_EPS = 1e-5                       # chiron/cnn.py:125 (epsilon=1e-5), :187
RNN_NORMAL, RNN_RNA = 0, 1
NORM_NONE, NORM_UNIQUE_MAD, NORM_FULL_MAD = 0, 1, 2
BN_POPULATION, BN_BATCH = 0, 1
CELL_LSTM, CELL_GRU = 0, 1
_HEADER = struct.Struct("<4s8i8i8i6iq")


@dataclass
class ModelConfig:
    n_blocks: int = 3
    channels: int = 256
    hidden: int = 100
    n_layers: int = 3
    n_class: int = 5
    rnn_layout: int = RNN_NORMAL
    branch1_bn_mask: int = 1
    k: List[int] = field(default_factory=lambda: [3, 3, 3])
    stride: List[int] = field(default_factory=lambda: [1, 1, 1])
    sig_norm: int = NORM_UNIQUE_MAD
    reverse_signal: int = 0
    bn_mode: int = BN_POPULATION
    cell_type: int = CELL_LSTM
    stem_k: int = 0
    stem_stride: int = 0

    def total_stride(self) -> int:
        s = self.stem_stride if self.stem_k else 1
        for v in self.stride[: self.n_blocks]:
            s *= v
        return s

    def out_len(self, L: int) -> int:
        """CNN output length for an L-sample window (TF 'SAME': ceil(T/stride) per strided conv)."""
        T = -(-L // self.stem_stride) if self.stem_k else L
        for b in range(self.n_blocks):
            T = -(-T // self.stride[b])
        return T

    def lstm_in(self, layer: int) -> int:
        if layer == 0:
            return self.channels
        return 2 * self.hidden if self.rnn_layout == RNN_NORMAL else self.hidden


def tensor_specs(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    C, H = cfg.channels, cfg.hidden
    specs: List[Tuple[str, Tuple[int, ...]]] = []

    def bn(prefix):
        for n in ("scale", "offset", "pop_mean", "pop_var"):
            specs.append(("%s_bn/%s" % (prefix, n), (C,)))

    if cfg.stem_k:
        specs.append(("conv_layer/conv1/weights", (cfg.stem_k, C)))
        bn("conv_layer/conv1")
    for b in range(cfg.n_blocks):
        cin = 1 if (b == 0 and not cfg.stem_k) else C
        p = "res_layer%d" % (b + 1)
        specs.append((p + "/branch1/conv1/weights", (cin, C)))
        if cfg.branch1_bn_mask >> b & 1:
            bn(p + "/branch1/conv1")
        specs.append((p + "/branch2/conv2a/weights", (cin, C)))
        bn(p + "/branch2/conv2a")
        specs.append((p + "/branch2/conv2b/weights", (cfg.k[b], C, C)))
        bn(p + "/branch2/conv2b")
        specs.append((p + "/branch2/conv2c/weights", (C, C)))
        bn(p + "/branch2/conv2c")
    for l in range(cfg.n_layers):
        for d in ("fw", "bw"):
            if cfg.cell_type == CELL_GRU:
                specs.append(("gru/%d/%s/gates/kernel" % (l, d), (cfg.lstm_in(l) + H, 2 * H)))
                specs.append(("gru/%d/%s/gates/bias" % (l, d), (2 * H,)))
                specs.append(("gru/%d/%s/candidate/kernel" % (l, d), (cfg.lstm_in(l) + H, H)))
                specs.append(("gru/%d/%s/candidate/bias" % (l, d), (H,)))
            else:
                specs.append(("lstm/%d/%s/kernel" % (l, d), (cfg.lstm_in(l) + H, 4 * H)))
                specs.append(("lstm/%d/%s/bias" % (l, d), (4 * H,)))
    specs.append(("rnn_fnn_layer/weights", (2, H)))
    specs.append(("rnn_fnn_layer/bias", (H,)))
    specs.append(("rnn_fnn_layer/weights_class", (H, cfg.n_class)))
    specs.append(("rnn_fnn_layer/bias_class", (cfg.n_class,)))
    return specs


def pack_blob(cfg: ModelConfig, tensors: Dict[str, np.ndarray]) -> bytes:
    chunks = []
    for name, shape in tensor_specs(cfg):
        t = np.ascontiguousarray(tensors[name], dtype="<f4")
        if tuple(t.shape) != tuple(shape):
            raise ValueError("tensor %s has shape %s, expected %s" % (name, t.shape, shape))
        chunks.append(t.reshape(-1))
    flat = np.concatenate(chunks)
    k = list(cfg.k) + [0] * (MAX_BLOCKS - len(cfg.k))
    s = list(cfg.stride) + [0] * (MAX_BLOCKS - len(cfg.stride))
    head = _HEADER.pack(MAGIC, 1, cfg.n_blocks, cfg.channels, cfg.hidden, cfg.n_layers, cfg.n_class, cfg.rnn_layout,
                        cfg.branch1_bn_mask, *k[:MAX_BLOCKS], *s[:MAX_BLOCKS], cfg.sig_norm, cfg.reverse_signal,
                        cfg.bn_mode, cfg.cell_type, cfg.stem_k, cfg.stem_stride, flat.size)
    return head + flat.tobytes()


def unpack_blob(blob: bytes) -> Tuple[ModelConfig, Dict[str, np.ndarray]]:
    vals = _HEADER.unpack_from(blob, 0)
    if vals[0] != MAGIC or vals[1] != 1:
        raise ValueError("not a CBW1 weight blob")
    n_blocks, channels, hidden, n_layers, n_class, rnn_layout, mask = vals[2:9]
    k = list(vals[9:17])[:n_blocks]
    s = list(vals[17:25])[:n_blocks]
    sig_norm, reverse_signal, bn_mode, cell_type, stem_k, stem_stride = vals[25:31]
    n_floats = vals[31]
    cfg = ModelConfig(n_blocks, channels, hidden, n_layers, n_class, rnn_layout, mask, k, s, sig_norm, reverse_signal,
                      bn_mode, cell_type, stem_k, stem_stride)
    flat = np.frombuffer(blob, dtype="<f4", count=n_floats, offset=_HEADER.size)
    tensors: Dict[str, np.ndarray] = {}
    pos = 0
    for name, shape in tensor_specs(cfg):
        n = int(np.prod(shape))
        tensors[name] = flat[pos:pos + n].reshape(shape)
        pos += n
    if pos != n_floats:
        raise ValueError("weight blob has %d floats, topology needs %d" % (n_floats, pos))
    return cfg, tensors


def random_tensors(cfg: ModelConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random-init weights of an arbitrary residual-stack + BiLSTM topology (there are shipped checkpoints only for
    DNA_default and RNA_default): the reference's initialisers where they matter for the scale of the activations --
    Xavier-normal convolutions (chiron/cnn.py:41-45), truncated-normal head (chiron/rnn.py:73-88), TF's default
    Glorot-uniform LSTM kernels -- and BN statistics of a plausible trained model.  Used by the parity tests and the
    configuration sweeps for topologies such as the 5-block ``rna_test`` (chiron/cnn.py:555-566)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in tensor_specs(cfg):
        leaf = name.rsplit("/", 1)[-1]
        if leaf == "scale":
            v = rng.uniform(0.6, 1.4, size=shape)
        elif leaf == "offset":
            v = rng.normal(0.0, 0.1, size=shape)
        elif leaf == "pop_mean":
            v = rng.normal(0.0, 0.1, size=shape)
        elif leaf == "pop_var":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == "kernel":
            lim = 3.0 * np.sqrt(6.0 / (shape[0] + shape[1]))     # 3x Glorot: gates that move with the input
            v = rng.uniform(-lim, lim, size=shape)
        elif leaf in ("bias", "bias_class"):
            v = rng.normal(0.0, 0.05, size=shape)
            if name.endswith("gates/bias"):
                v = v + 1.0                      # GRUCell's gate bias initialiser
        elif name == "rnn_fnn_layer/weights":
            v = rng.normal(0.0, np.sqrt(2.0 / (2 * cfg.hidden)), size=shape) + 0.5
        elif name == "rnn_fnn_layer/weights_class":
            v = rng.normal(0.0, 8.0 * np.sqrt(2.0 / cfg.hidden), size=shape)    # logits of a trained model's magnitude
        else:                                   # convolution weights [.., fan_in, C]
            fan_in = int(np.prod(shape[:-1]))
            from_signal = fan_in == 1 or name.startswith("conv_layer/")        # reads the raw one-channel signal
            v = rng.normal(0.0, np.sqrt(2.0 / (fan_in + shape[-1])) * (3.0 if from_signal else 1.4), size=shape)
        out[name] = v.astype(np.float32)
    return out


def header_size() -> int:
    return _HEADER.size


_WEIGHT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")


def bundled_blob_path(model: str) -> str:
    """Path of a converted bundled model ("DNA_default" / "RNA_default")."""
    return os.path.join(_WEIGHT_DIR, model + ".cbw")


def resolve_model(model: str) -> str:
    """Map the reference's ``-m <model folder>`` argument (chiron/entry.py:75) to a packed blob.

    Accepts a ``.cbw`` file, a folder containing one, a folder name of a bundled model, or a TF checkpoint folder
    (converted on the fly with the dependency-free bundle reader)."""
    if os.path.isfile(model):
        return model
    base = os.path.basename(os.path.normpath(model))
    if os.path.isdir(model):
        for fn in sorted(os.listdir(model)):
            if fn.endswith(".cbw"):
                return os.path.join(model, fn)
        if os.path.exists(os.path.join(model, "checkpoint")):
            return model                        # a TF checkpoint folder: load_model converts it in memory
    if os.path.exists(bundled_blob_path(base)):
        return bundled_blob_path(base)
    raise FileNotFoundError("cannot resolve model %r to a CBW1 weight blob" % model)


def load_model(model: str) -> Tuple[ModelConfig, Dict[str, np.ndarray], bytes]:
    path = resolve_model(model)
    if os.path.isdir(path):                     # TF checkpoint folder: converted on the fly, nothing written to disk
        from .convert_weights import convert_checkpoint_dir
        blob = convert_checkpoint_dir(path)
    else:
        with open(path, "rb") as f:
            blob = f.read()
    cfg, tensors = unpack_blob(blob)
    return cfg, tensors, blob
