"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path (``chiron_b200/``).

A numpy restatement of the reference's basecalling inference path (`chiron call` ->
``chiron_eval.evaluation``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this file, and only as the checker / CPU baseline.

The arithmetic of the reference lives in TensorFlow 1.15 (setup.py:28), which is absent from /root/reference and
cannot be installed here; the functions below restate the published semantics of the TF ops at the reference's
call sites (cited per function) -- see SURVEY.md App. A.  PARITY PINNING: this oracle is pinned by the reference's own
golden fixtures (chiron/example_data/DNA/output, copied to tests/golden/): raw/read1.signal -> segments/read1.fastq
(161/161 windows, beam 30) and raw/read3.signal -> segments/read3.fastq (319/319), segments/readN.fastq ->
result/readN.fastq for N=1..5 and the quality string of result/read1.fastq (tests/test_oracle_golden.py).
RNA_default has no golden outputs in the reference: its stride-5/k=13 block and MultiRNNCell layout are
"parity unpinned".

Everything takes a ``dtype`` (np.float32 = the reference's compute type; np.float64 = noise-floor reference).
"""
from __future__ import annotations

import difflib
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BN_EPS = 1e-5
MAD_SCALE = 0.6744897501960817            # statsmodels.robust.mad normalisation constant (Phi^-1(3/4))
NEG_INF = float("-inf")


# ----------------------------------------------------------------------------------------------------------------------
# A.1 / A.2  signal -> normalised windows
# ----------------------------------------------------------------------------------------------------------------------
def read_signal_text(path: str) -> np.ndarray:
    """chiron/chiron_input.py:527-532 -- whitespace separated numbers, parsed as float32."""
    with open(path) as f:
        return np.asarray(f.read().split(), dtype=np.float32)


def normalize_signal(signal: np.ndarray, mode: int = 1) -> np.ndarray:
    """chiron/chiron_input.py:548-554 (MEDIAN branch): statistics over the *unique* values (mode 1) as
    read_signal_fast5 does, or over the full signal (mode 2) as read_signal does (:535-538); mode 0 = none.
    statsmodels.robust.mad(a) = median(|a - median(a)|) / 0.67448975...  Arithmetic in float64, result float32."""
    s = np.asarray(signal, dtype=np.float64)
    if mode == 0 or s.size == 0:
        return s.astype(np.float32)
    ref = np.unique(s) if mode == 1 else s
    med = np.median(ref)
    mad = np.median(np.abs(ref - med)) / MAD_SCALE
    return ((s - med) / mad).astype(np.float32)


def make_windows(signal: np.ndarray, seg_len: int, jump: int, start: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """chiron/chiron_input.py:279-286 + padding() :681-692 -- windows at 0,jump,2*jump..<n, zero padded."""
    sig = np.asarray(signal, dtype=np.float32)[start:]
    n = sig.shape[0]
    starts = list(range(0, n, jump))
    x = np.zeros((len(starts), seg_len), dtype=np.float32)
    lens = np.zeros(len(starts), dtype=np.int32)
    for i, s in enumerate(starts):
        w = sig[s:s + seg_len]
        x[i, :w.shape[0]] = w
        lens[i] = w.shape[0]
    return x, lens


def seq_len_out(lens: np.ndarray, ratio: float) -> np.ndarray:
    """chiron/chiron_eval.py:337 -- np.round(seq_len/ratio).astype(int32) (round half to even)."""
    return np.round(np.asarray(lens, dtype=np.float64) / ratio).astype(np.int32)


# ----------------------------------------------------------------------------------------------------------------------
# A.3 / A.4  CNN
# ----------------------------------------------------------------------------------------------------------------------
def _bn(x, t, prefix, dtype, bn_mode=0, stats_out=None):
    """bn_mode 0: tf.nn.batch_normalization with population statistics (chiron/cnn.py:160-161, the graph of the shipped
    checkpoints).  bn_mode 1: HEAD's simple_global_bn (chiron/cnn.py:166-188): tf.nn.moments(inp, [0, 1, 2]) of THIS
    batch -- every frame of every window, zero padding included -- mean = E[x], var = E[(x - mean)^2], even at
    inference.  Either way  inv = scale*rsqrt(var+eps); y = x*inv + (offset - mean*inv).
    ``stats_out`` (dict) receives the (mean, var) used, keyed by ``prefix``.
    Parity of mode 1 is UNPINNED: no checkpoint trained at HEAD ships with the reference (SURVEY.md finding 2)."""
    scale = t[prefix + "_bn/scale"].astype(dtype)
    offset = t[prefix + "_bn/offset"].astype(dtype)
    if bn_mode == 0:
        mean = t[prefix + "_bn/pop_mean"].astype(dtype)
        var = t[prefix + "_bn/pop_var"].astype(dtype)
    else:
        flat = x.reshape(-1, x.shape[-1])
        mean = flat.mean(axis=0, dtype=dtype)
        var = np.square(flat - mean).mean(axis=0, dtype=dtype)
    if stats_out is not None:
        stats_out[prefix] = (mean, var)
    inv = scale * (dtype(1.0) / np.sqrt(var + dtype(BN_EPS)))
    return x * inv + (offset - mean * inv)


def _conv_same(x, w, stride):
    """tf.nn.conv2d(NHWC, 'SAME') along time (chiron/cnn.py:60-64): cross-correlation, x [B,T,Cin], w [k,Cin,Cout]."""
    B, T, _ = x.shape
    k = w.shape[0]
    t_out = -(-T // stride)
    pad = max((t_out - 1) * stride + k - T, 0)
    left = pad // 2
    xp = np.zeros((B, T + pad, x.shape[2]), dtype=x.dtype)
    xp[:, left:left + T] = x
    out = None
    span = (t_out - 1) * stride + 1
    for j in range(k):
        term = xp[:, j:j + span:stride] @ w[j]
        out = term if out is None else out + term
    return out


def cnn_forward(x: np.ndarray, cfg, t: Dict[str, np.ndarray], dtype=np.float32, bn_mode=None,
                stats_out=None) -> np.ndarray:
    """getcnnfeature -> DNA_model1 / rna_test / ... -> residual_layer (chiron/cnn.py:334-371, 380-389, 555-566,
    234-262).  x [B,L] -> [B,T,C].  ``bn_mode`` None = the model's own (cfg.bn_mode); see ``_bn``."""
    if bn_mode is None:
        bn_mode = getattr(cfg, "bn_mode", 0)
    net = np.asarray(x, dtype=dtype)[:, :, None]
    if getattr(cfg, "stem_k", 0):        # RNA_model2 / RNA_model3 (chiron/cnn.py:454-476): conv_layer [1,k,1,C], stride, BN, ReLU
        w = t["conv_layer/conv1/weights"].astype(dtype)[:, None, :]
        net = np.maximum(_bn(_conv_same(net, w, cfg.stem_stride), t, "conv_layer/conv1", dtype, bn_mode, stats_out), 0)
    for b in range(cfg.n_blocks):
        p = "res_layer%d" % (b + 1)
        s = cfg.stride[b]
        w1 = t[p + "/branch1/conv1/weights"].astype(dtype)[None]
        b1 = _conv_same(net, w1, s)
        if cfg.branch1_bn_mask >> b & 1:
            b1 = _bn(b1, t, p + "/branch1/conv1", dtype, bn_mode, stats_out)
        a = _conv_same(net, t[p + "/branch2/conv2a/weights"].astype(dtype)[None], 1)
        a = np.maximum(_bn(a, t, p + "/branch2/conv2a", dtype, bn_mode, stats_out), 0)
        bb = _conv_same(a, t[p + "/branch2/conv2b/weights"].astype(dtype), s)
        bb = np.maximum(_bn(bb, t, p + "/branch2/conv2b", dtype, bn_mode, stats_out), 0)
        c = _conv_same(bb, t[p + "/branch2/conv2c/weights"].astype(dtype)[None], 1)
        c = _bn(c, t, p + "/branch2/conv2c", dtype, bn_mode, stats_out)
        net = np.maximum(b1 + c, 0)
    return net


# ----------------------------------------------------------------------------------------------------------------------
# A.5  BiLSTM   (TF LSTMCell: gates i,j,f,o; forget_bias 1.0; dynamic_rnn sequence_length semantics)
# ----------------------------------------------------------------------------------------------------------------------
def _sigmoid(x):
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


def lstm_direction(x: np.ndarray, lens: np.ndarray, kernel: np.ndarray, bias: np.ndarray, reverse: bool,
                   dtype=np.float32) -> np.ndarray:
    """One LSTMCell run by dynamic_rnn (chiron/rnn.py:49-50,64): z = concat[x_t,h]@K + b; i,j,f,o = split(z);
    c' = sigmoid(f+1)*c + sigmoid(i)*tanh(j); h' = sigmoid(o)*tanh(c').  For t >= len the output is zero and the
    state is frozen.  ``reverse`` = array_ops.reverse_sequence on the first len frames (backward direction)."""
    B, T, D = x.shape
    H = bias.shape[0] // 4
    kernel = kernel.astype(dtype)
    bias = bias.astype(dtype)
    wx, wh = kernel[:D], kernel[D:]
    pre = x.reshape(B * T, D) @ wx
    pre = pre.reshape(B, T, 4 * H)
    h = np.zeros((B, H), dtype=dtype)
    c = np.zeros((B, H), dtype=dtype)
    out = np.zeros((B, T, H), dtype=dtype)
    one = dtype(1.0)
    lens = np.asarray(lens)
    max_len = int(lens.max()) if B else 0
    for step in range(max_len):
        active = step < lens
        if reverse:
            tt = np.where(active, lens - 1 - step, 0)
        else:
            tt = np.full(B, step)
        z = pre[np.arange(B), tt] + h @ wh + bias
        i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        c_new = _sigmoid(f + one) * c + _sigmoid(i) * np.tanh(j)
        h_new = _sigmoid(o) * np.tanh(c_new)
        a = active[:, None]
        c = np.where(a, c_new, c).astype(dtype)
        h = np.where(a, h_new, h).astype(dtype)
        rows = np.nonzero(active)[0]
        out[rows, tt[rows]] = h_new[rows]
    return out


def gru_direction(x: np.ndarray, lens: np.ndarray, gate_kernel: np.ndarray, gate_bias: np.ndarray,
                  cand_kernel: np.ndarray, cand_bias: np.ndarray, reverse: bool, dtype=np.float32) -> np.ndarray:
    """One GRUCell run by dynamic_rnn (chiron/rnn.py:51-53,129-131; TF 1.15 rnn_cell_impl.GRUCell.call):
    r, u = split(sigmoid(concat[x_t, h] @ Kg + bg));  c = tanh(concat[x_t, r*h] @ Kc + bc);  h' = u*h + (1-u)*c.
    sequence_length / reverse semantics as in ``lstm_direction``.  PARITY UNPINNED: no GRU checkpoint ships."""
    B, T, D = x.shape
    H = cand_bias.shape[0]
    gate_kernel, cand_kernel = gate_kernel.astype(dtype), cand_kernel.astype(dtype)
    gate_bias, cand_bias = gate_bias.astype(dtype), cand_bias.astype(dtype)
    pre_g = (x.reshape(B * T, D) @ gate_kernel[:D]).reshape(B, T, 2 * H)
    pre_c = (x.reshape(B * T, D) @ cand_kernel[:D]).reshape(B, T, H)
    wg_h, wc_h = gate_kernel[D:], cand_kernel[D:]
    h = np.zeros((B, H), dtype=dtype)
    out = np.zeros((B, T, H), dtype=dtype)
    one = dtype(1.0)
    lens = np.asarray(lens)
    max_len = int(lens.max()) if B else 0
    rows_all = np.arange(B)
    for step in range(max_len):
        active = step < lens
        tt = np.where(active, lens - 1 - step, 0) if reverse else np.full(B, step)
        g = _sigmoid(pre_g[rows_all, tt] + h @ wg_h + gate_bias)
        r, u = g[:, :H], g[:, H:]
        c = np.tanh(pre_c[rows_all, tt] + (r * h) @ wc_h + cand_bias)
        h_new = u * h + (one - u) * c
        h = np.where(active[:, None], h_new, h).astype(dtype)
        rows = np.nonzero(active)[0]
        out[rows, tt[rows]] = h_new[rows]
    return out


def _cell_direction(x, lens, cfg, t, l, d, reverse, dtype):
    if getattr(cfg, "cell_type", 0) == 1:
        p = "gru/%d/%s/" % (l, d)
        return gru_direction(x, lens, t[p + "gates/kernel"], t[p + "gates/bias"], t[p + "candidate/kernel"],
                             t[p + "candidate/bias"], reverse, dtype)
    return lstm_direction(x, lens, t["lstm/%d/%s/kernel" % (l, d)], t["lstm/%d/%s/bias" % (l, d)], reverse, dtype)


def rnn_forward(fea: np.ndarray, lens: np.ndarray, cfg, t: Dict[str, np.ndarray], dtype=np.float32) -> np.ndarray:
    """rnn_layers (stack_bidirectional_dynamic_rnn, chiron/rnn.py:20-64) or rnn_layers_rna (MultiRNNCell per direction
    + bidirectional_dynamic_rnn, chiron/rnn.py:99-145).  Returns lasth [B,T,2H]."""
    fea = np.asarray(fea, dtype=dtype)
    if cfg.rnn_layout == 0:
        x = fea
        for l in range(cfg.n_layers):
            fw = _cell_direction(x, lens, cfg, t, l, "fw", False, dtype)
            bw = _cell_direction(x, lens, cfg, t, l, "bw", True, dtype)
            x = np.concatenate([fw, bw], axis=2)
        return x
    outs = []
    for d, rev in (("fw", False), ("bw", True)):
        x = fea
        for l in range(cfg.n_layers):
            # inside MultiRNNCell every layer sees the same (possibly reversed) time order, so running each layer as
            # its own reversed pass is identical to reversing once around the stack.
            x = _cell_direction(x, lens, cfg, t, l, d, rev, dtype)
        outs.append(x)
    return np.concatenate(outs, axis=2)


# ----------------------------------------------------------------------------------------------------------------------
# A.6  head, path_prob
# ----------------------------------------------------------------------------------------------------------------------
def head_forward(lasth: np.ndarray, cfg, t: Dict[str, np.ndarray], dtype=np.float32) -> np.ndarray:
    """chiron/rnn.py:89-96: h2 = fw*W[0] + bw*W[1] + bias ; logits = h2 @ Wc + bc."""
    B, T, _ = lasth.shape
    H = cfg.hidden
    w = t["rnn_fnn_layer/weights"].astype(dtype)
    h2 = lasth[:, :, :H] * w[0] + lasth[:, :, H:] * w[1]
    h2 = h2 + t["rnn_fnn_layer/bias"].astype(dtype)
    logits = h2.reshape(B * T, H) @ t["rnn_fnn_layer/weights_class"].astype(dtype)
    logits = logits + t["rnn_fnn_layer/bias_class"].astype(dtype)
    return logits.reshape(B, T, cfg.n_class)


def inference(x: np.ndarray, lens_out: np.ndarray, cfg, t, dtype=np.float32, bn_mode=None) -> np.ndarray:
    """chiron/chiron_model.py:134-172: CNN -> RNN -> logits[B,T,n_class].  ``lens_out`` already divided by ratio."""
    fea = cnn_forward(x, cfg, t, dtype, bn_mode)
    lasth = rnn_forward(fea, lens_out, cfg, t, dtype)
    return head_forward(lasth, cfg, t, dtype)


def path_prob(logits: np.ndarray) -> np.ndarray:
    """chiron/chiron_eval.py:116-136: mean over ALL T frames of (top1 - top2) logits.  Returns [B] float32."""
    s = np.sort(np.asarray(logits, dtype=np.float32), axis=2)
    return (s[:, :, -1] - s[:, :, -2]).mean(axis=1, dtype=np.float32)


# ----------------------------------------------------------------------------------------------------------------------
# A.7a greedy / A.7b beam search   (TF CTC ops; blank = last class)
# ----------------------------------------------------------------------------------------------------------------------
def ctc_greedy(logits: np.ndarray, lens: np.ndarray) -> List[List[int]]:
    """tf.nn.ctc_greedy_decoder(merge_repeated=True) (chiron/chiron_eval.py:486-487)."""
    B, T, C = logits.shape
    blank = C - 1
    am = np.argmax(logits, axis=2)          # first maximum wins, like Eigen's maxCoeff(&idx)
    out = []
    for b in range(B):
        seq, prev = [], -1
        for tt in range(int(lens[b])):
            m = int(am[b, tt])
            if m != blank and m != prev:
                seq.append(m)
            prev = m
        out.append(seq)
    return out


def _lse(a: np.float32, b: np.float32) -> np.float32:
    if a == NEG_INF and b == NEG_INF:
        return np.float32(NEG_INF)
    if a > b:
        return np.float32(a + np.log1p(np.exp(np.float32(b - a), dtype=np.float32), dtype=np.float32))
    return np.float32(b + np.log1p(np.exp(np.float32(a - b), dtype=np.float32), dtype=np.float32))


class _Beam:
    __slots__ = ("parent", "label", "children", "o_total", "o_blank", "o_label", "n_total", "n_blank", "n_label")

    def __init__(self, parent, label):
        self.parent = parent
        self.label = label
        self.children = None
        ninf = np.float32(NEG_INF)
        self.o_total = self.o_blank = self.o_label = ninf
        self.n_total = self.n_blank = self.n_label = ninf

    def active(self):
        return self.n_total != NEG_INF

    def reset_new(self):
        self.n_total = self.n_blank = self.n_label = np.float32(NEG_INF)

    def reset_old(self):
        self.o_total = self.o_blank = self.o_label = np.float32(NEG_INF)


def ctc_beam_search_one(logits: np.ndarray, length: int, beam_width: int) -> List[int]:
    """tensorflow/core/util/ctc/ctc_beam_search.h (TF 1.15) CTCBeamSearchDecoder::Step / TopPaths with
    merge_repeated=False, top_paths=1 -- the decoder behind tf.nn.ctc_beam_search_decoder
    (chiron/chiron_eval.py:489-492).  float32 log-space arithmetic; `leaves` is a TopN(beam_width) by newp.total."""
    C = logits.shape[1]
    blank = C - 1
    f32 = np.float32
    root = _Beam(None, -1)
    root.n_total = f32(0.0)
    root.n_blank = f32(0.0)
    leaves: List[_Beam] = [root]
    for tt in range(length):
        row = logits[tt].astype(np.float32)
        inp = row - row.max()
        # leaves_.Extract(): descending newp.total (stable for ties: insertion order)
        branches = sorted(leaves, key=lambda e: -float(e.n_total))
        leaves = []
        for b in branches:
            b.o_total, b.o_blank, b.o_label = b.n_total, b.n_blank, b.n_label
        for b in branches:
            if b.parent is not None:
                if b.parent.active():
                    prev = b.parent.o_blank if b.label == b.parent.label else b.parent.o_total
                    b.n_label = _lse(b.n_label, prev)
                b.n_label = f32(b.n_label + inp[b.label])
            b.n_blank = f32(b.o_total + inp[blank])
            b.n_total = _lse(b.n_blank, b.n_label)
            leaves.append(b)

        def bottom():
            return min(leaves, key=lambda e: float(e.n_total))

        def is_candidate(total):
            return total > NEG_INF and (len(leaves) < beam_width or total > bottom().n_total)

        for b in branches:
            if not is_candidate(b.o_total):
                continue
            if b.children is None:
                b.children = [_Beam(b, lab) for lab in range(C - 1)]
            for c in b.children:
                if c.active():
                    continue
                c.n_blank = f32(NEG_INF)
                prev = b.o_blank if c.label == b.label else b.o_total
                c.n_label = f32(inp[c.label] + prev)
                c.n_total = c.n_label
                if is_candidate(c.n_total):
                    if len(leaves) == beam_width:
                        bot = bottom()
                        leaves.remove(bot)
                        bot.reset_new()
                    leaves.append(c)
                else:
                    c.reset_old()
                    c.reset_new()
    best = max(leaves, key=lambda e: float(e.n_total))
    seq = []
    while best.parent is not None:
        seq.append(best.label)
        best = best.parent
    return seq[::-1]


def ctc_beam_search(logits: np.ndarray, lens: np.ndarray, beam_width: int) -> List[List[int]]:
    return [ctc_beam_search_one(logits[b], int(lens[b]), beam_width) for b in range(logits.shape[0])]


# ----------------------------------------------------------------------------------------------------------------------
# A.8  assembly   (chiron/utils/easy_assembler.py)
# ----------------------------------------------------------------------------------------------------------------------
BASES = "ACGT"


def index2base(read: Sequence[int]) -> str:
    """chiron/chiron_eval.py:100-113."""
    return "".join(BASES[int(x)] for x in read)


def get_assembler_kernal(jump: int, segment_len: int) -> str:
    """chiron/chiron_eval.py:138-150."""
    assembler = "simple"
    if jump > 0.9 * segment_len:
        assembler = "glue"
    if jump >= segment_len:
        assembler = "stick"
    return assembler


def stick_kernal(bpread: str, prev_bpread: str) -> int:
    """chiron/utils/easy_assembler.py:296-300."""
    return len(prev_bpread)


def glue_kernal(bpread: str, prev_bpread: str) -> int:
    """chiron/utils/easy_assembler.py:276-294."""
    prev_n = len(prev_bpread)
    n = len(bpread)
    max_overlap = min(math.floor(0.1 * prev_n), n)
    best_i, best_score = 0, 0
    for i in range(1, max_overlap):
        hits = sum(1 for a, b in zip(bpread[:i], prev_bpread[-i:]) if a == b)
        score = 2 * hits - i
        if score > best_score:
            best_i, best_score = i, score
    return prev_n - best_i


def simple_assembly_kernal(bpread: str, prev_bpread: str, error_rate: float, jump_step_ratio: float) -> int:
    """chiron/utils/easy_assembler.py:212-250 (difflib.SequenceMatcher with default autojunk)."""
    back_ratio = 6.5 * 10e-4
    p_same = 1 - 2 * error_rate + 26 / 25 * (error_rate ** 2)
    ns: Dict[int, int] = {}
    N = len(bpread)
    for block in difflib.SequenceMatcher(a=bpread, b=prev_bpread).get_matching_blocks():
        offset = block[1] - block[0]
        ns[offset] = ns.get(offset, 0) + block[2]
    log_px: Dict[int, float] = {}
    for key in ns:
        kk = -key if key < 0 else key
        rate = back_ratio * N * jump_step_ratio if key < 0 else N * jump_step_ratio
        log_px[key] = kk * np.log(rate) - sum(np.log(x + 1) for x in range(kk)) + ns[key] * np.log(p_same / 0.25)
    return max(log_px.keys(), key=lambda x: log_px[x])


def pair_displacement(bpread: str, prev_bpread: str, kernal: str, jump_step_ratio: float,
                      error_rate: float = 0.2) -> int:
    if kernal == "simple":
        return simple_assembly_kernal(bpread, prev_bpread, error_rate, jump_step_ratio)
    if kernal == "glue":
        return glue_kernal(bpread, prev_bpread)
    if kernal == "stick":
        return stick_kernal(bpread, prev_bpread)
    raise ValueError("unsupported assembly kernel %r" % kernal)


def simple_assembly_qs(bpreads: Sequence[str], qs_list: Optional[Sequence[float]], jump_step_ratio: float,
                       error_rate: float = 0.2, kernal: str = "simple"):
    """simple_assembly / simple_assembly_qs + add_count(_qs) (chiron/utils/easy_assembler.py:302-335,381-442).
    Returns (consensus[4,length] f64, consensus_qs[4,length] f64, pos[n] int) -- ``pos`` are the window start
    coordinates; like the reference, window 0 does not update ``length``."""
    census_len = 1000
    cons = np.zeros((4, census_len))
    cons_qs = np.zeros((4, census_len))
    pos, length = 0, 0
    positions = []
    base_idx = {"A": 0, "C": 1, "G": 2, "T": 3}

    def add(start, seg, q):
        if start < 0:
            seg = seg[-start:]
            start = 0
        for i, ch in enumerate(seg):
            cons[base_idx[ch], start + i] += 1
            cons_qs[base_idx[ch], start + i] += q

    for indx, bpread in enumerate(bpreads):
        q = float(qs_list[indx]) if qs_list is not None else 0.0
        if indx == 0:
            add(0, bpread, q)
            positions.append(0)
            continue
        disp = pair_displacement(bpread, bpreads[indx - 1], kernal, jump_step_ratio, error_rate)
        if disp + pos + len(bpread) > census_len:
            cons = np.pad(cons, ((0, 0), (0, 1000)))
            cons_qs = np.pad(cons_qs, ((0, 0), (0, 1000)))
            census_len += 1000
        add(pos + disp, bpread, q)
        pos += disp
        positions.append(pos)
        length = max(length, pos + len(bpread))
    return cons[:, :length], cons_qs[:, :length], np.asarray(positions, dtype=np.int64)


def qs_string(consensus: np.ndarray, consensus_qs: np.ndarray) -> str:
    """chiron/chiron_eval.py:152-174 (phred+33).  The reference calls np.argsort(consensus, axis=0) with the default
    kind; on the numpy of its era (<= 1.16) a 4-element axis is insertion-sorted, i.e. stable, which decides which
    row's quality sum is used when the two highest counts tie.  Modern numpy's SIMD argsort is not stable, so the
    restatement asks for kind="stable" explicitly (the golden quality string of read1 pins this)."""
    if consensus.shape[1] == 0:
        return ""
    sort_ind = np.argsort(consensus, axis=0, kind="stable")
    L = consensus.shape[1]
    sc = consensus[sort_ind, np.arange(L)[np.newaxis, :]]
    sq = consensus_qs[sort_ind, np.arange(L)[np.newaxis, :]]
    with np.errstate(divide="ignore", invalid="ignore"):
        quality = 10 * (np.log10((sc[3, :] + 1) / (sc[2, :] + 1))) + sq[3, :] / sc[3, :] / np.log(10)
    return "".join(chr(int(x) + 33) for x in quality.astype(int))


# ----------------------------------------------------------------------------------------------------------------------
# whole read:  signal -> windows -> logits -> decode -> assembly   (chiron/chiron_eval.py:378-462, true window order)
# ----------------------------------------------------------------------------------------------------------------------
def basecall_signal(raw_signal: np.ndarray, cfg, t, seg_len: int, jump: int, beam: int = 0, start: int = 0,
                    batch: int = 0, dtype=np.float32, with_qs: bool = True):
    sig = np.asarray(raw_signal)
    if cfg.reverse_signal:
        sig = sig[::-1]
    x, lens = make_windows(normalize_signal(sig, cfg.sig_norm), seg_len, jump, start)
    T = cfg.out_len(seg_len)
    lens_o = seq_len_out(lens, seg_len / T)
    n = x.shape[0]
    step = batch if batch > 0 else max(n, 1)
    logits = np.concatenate([inference(x[i:i + step], lens_o[i:i + step], cfg, t, dtype)
                             for i in range(0, n, step)]) if n else np.zeros((0, T, cfg.n_class), np.float32)
    logits32 = logits.astype(np.float32)
    paths = ctc_greedy(logits32, lens_o) if beam == 0 else ctc_beam_search(logits32, lens_o, beam)
    pp = path_prob(logits32)
    keep = [i for i, p in enumerate(paths) if len(p) > 0]     # sparse2dense drops empty rows (chiron_eval.py:56-66)
    bpreads = [index2base(paths[i]) for i in keep]
    qs_list = [pp[i] for i in keep]
    kernal = get_assembler_kernal(jump, seg_len)
    cons, cons_qs, pos = simple_assembly_qs(bpreads, qs_list if with_qs else None, jump / seg_len, kernal=kernal)
    seq = index2base(np.argmax(cons, axis=0)) if cons.shape[1] else ""
    qual = qs_string(cons, cons_qs) if with_qs else None
    return {"logits": logits, "lens": lens_o, "paths": paths, "path_prob": pp, "segments": bpreads,
            "consensus": seq, "qual": qual, "pos": pos, "x": x}


# ----------------------------------------------------------------------------------------------------------------------
# fast path of the same decoders: oracle/ctc_oracle.c through ctypes (identical algorithm and tie rules)
# ----------------------------------------------------------------------------------------------------------------------
_CLIB = None


def _clib():
    global _CLIB
    if _CLIB is None:
        import ctypes
        from . import build as _build
        _CLIB = ctypes.CDLL(_build.build())
    return _CLIB


def ctc_decode_c(logits: np.ndarray, lens: np.ndarray, beam_width: int = 0) -> List[List[int]]:
    """Greedy (beam_width=0) or TF-style beam search through the C restatement."""
    import ctypes
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    B, T, C = lg.shape
    ln = np.ascontiguousarray(lens, dtype=np.int32)
    out = np.zeros((B, max(T, 1)), dtype=np.int32)
    out_len = np.zeros(B, dtype=np.int32)
    lib = _clib()
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int)
    if beam_width > 0:
        lib.oracle_ctc_beam_batch(lg.ctypes.data_as(fp), B, T, C, ln.ctypes.data_as(ip), int(beam_width),
                                  out.ctypes.data_as(ip), out_len.ctypes.data_as(ip))
    else:
        lib.oracle_ctc_greedy_batch(lg.ctypes.data_as(fp), B, T, C, ln.ctypes.data_as(ip),
                                    out.ctypes.data_as(ip), out_len.ctypes.data_as(ip))
    if (out_len < 0).any():
        raise MemoryError("oracle_ctc_beam failed")
    return [out[b, :out_len[b]].tolist() for b in range(B)]


def ctc_beam_scores_c(logits: np.ndarray, lens: np.ndarray, beam_width: int):
    """Beam search through the C restatement, with the top path's score (newp.total of the best beam: the `log_prob`
    output of tf.nn.ctc_beam_search_decoder as chiron/export_test.py:36-40 exposes it).  Returns (paths, scores[B])."""
    import ctypes
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    B, T, C = lg.shape
    ln = np.ascontiguousarray(lens, dtype=np.int32)
    out = np.zeros((B, max(T, 1)), dtype=np.int32)
    out_len = np.zeros(B, dtype=np.int32)
    scores = np.zeros(B, dtype=np.float32)
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
    _clib().oracle_ctc_beam_batch_scored(lg.ctypes.data_as(fp), B, T, C, ln.ctypes.data_as(ip), int(beam_width),
                                         out.ctypes.data_as(ip), out_len.ctypes.data_as(ip), scores.ctypes.data_as(fp))
    if (out_len < 0).any():
        raise MemoryError("oracle_ctc_beam failed")
    return [out[b, :out_len[b]].tolist() for b in range(B)], scores
