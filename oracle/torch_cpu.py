"""CPU BASELINE -- TEST / BENCH INFRASTRUCTURE ONLY.  Never imported by the product path (``chiron_b200/``).

The reference's CPU inference path (`chiron call` on TensorFlow 1.15's Eigen/MKL CPU kernels, chiron/chiron_eval.py:304-492)
restated on torch's CPU kernels (oneDNN convolutions, the fused ATen LSTM, MKL GEMM), because TF 1.15 cannot be installed
here: the best-effort stand-in for "the reference's own CPU implementation with all the host threads it can use" that
``bench.py --impl reference`` and its ``cpu_baseline`` leg time (torch-CPU, per-stage split, 1-thread
figure).  Same arithmetic as oracle/chiron_oracle.py (float32, population BatchNorm, TF LSTMCell gate order and
forget_bias, stack_bidirectional_dynamic_rnn / MultiRNNCell layouts) -- tests/test_torch_cpu_baseline.py pins it to that
oracle -- with the throughput a vectorised, multi-threaded CPU library gives; the numpy oracle's per-time-step Python loop
would understate the CPU by an order of magnitude.

Only LSTM cells and block stacks without a stem convolution are covered (the two shipped models and `rna_test`)."""
from __future__ import annotations

import time
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


class TorchCpuModel:
    def __init__(self, cfg, t: Dict[str, np.ndarray]):
        if getattr(cfg, "cell_type", 0) != 0 or getattr(cfg, "stem_k", 0) or getattr(cfg, "bn_mode", 0) != 0:
            raise ValueError("torch CPU baseline covers LSTM models with population BatchNorm and no stem convolution")
        self.cfg = cfg
        f = lambda a: torch.from_numpy(np.array(a, dtype=np.float32, order="C"))

        def bn(prefix):                    # tf.nn.batch_normalization with population statistics (chiron/cnn.py:160-161)
            inv = t[prefix + "_bn/scale"].astype(np.float32) / np.sqrt(t[prefix + "_bn/pop_var"].astype(np.float32) + np.float32(BN_EPS))
            return f(inv), f(t[prefix + "_bn/offset"].astype(np.float32) - t[prefix + "_bn/pop_mean"].astype(np.float32) * inv)

        self.blocks = []
        for b in range(cfg.n_blocks):
            p = "res_layer%d" % (b + 1)
            blk = {"stride": cfg.stride[b], "k": cfg.k[b]}
            # conv weights [k, Cin, Cout] (TF HWIO with H = 1) -> torch conv1d [Cout, Cin, k] (cross-correlation in both)
            w1 = t[p + "/branch1/conv1/weights"][None] if t[p + "/branch1/conv1/weights"].ndim == 2 else t[p + "/branch1/conv1/weights"]
            blk["w1"] = f(np.transpose(w1, (2, 1, 0)))
            blk["bn1"] = bn(p + "/branch1/conv1") if (cfg.branch1_bn_mask >> b) & 1 else None
            w2a = t[p + "/branch2/conv2a/weights"]
            blk["w2a"] = f(np.transpose(w2a[None] if w2a.ndim == 2 else w2a, (2, 1, 0)))
            blk["bn2a"] = bn(p + "/branch2/conv2a")
            blk["w2b"] = f(np.transpose(t[p + "/branch2/conv2b/weights"], (2, 1, 0)))
            blk["bn2b"] = bn(p + "/branch2/conv2b")
            w2c = t[p + "/branch2/conv2c/weights"]
            blk["w2c"] = f(np.transpose(w2c[None] if w2c.ndim == 2 else w2c, (2, 1, 0)))
            blk["bn2c"] = bn(p + "/branch2/conv2c")
            self.blocks.append(blk)
        H = cfg.hidden
        self.lstms = []                    # rnn_layout 0: one bidirectional stack; 1: one unidirectional stack per direction

        def load_dir(lstm, l, d, suffix):
            # TF LSTMCell kernel [in+H, 4H], gate columns i,j,f,o, forget_bias 1.0 added to f at run time
            # torch: weight_ih [4H, in], weight_hh [4H, H], gate rows i,f,g,o; both biases are added
            k, bias = t["lstm/%d/%s/kernel" % (l, d)].astype(np.float32), t["lstm/%d/%s/bias" % (l, d)].astype(np.float32).copy()
            d_in = k.shape[0] - H
            order = [0, 2, 1, 3]           # torch (i, f, g, o)  <-  TF (i, j, f, o)
            cols = np.concatenate([np.arange(g * H, (g + 1) * H) for g in order])
            bias[2 * H:3 * H] += 1.0
            with torch.no_grad():
                getattr(lstm, "weight_ih_l%d%s" % (l, suffix)).copy_(f(k[:d_in, cols].T))
                getattr(lstm, "weight_hh_l%d%s" % (l, suffix)).copy_(f(k[d_in:, cols].T))
                getattr(lstm, "bias_ih_l%d%s" % (l, suffix)).copy_(f(bias[cols]))
                getattr(lstm, "bias_hh_l%d%s" % (l, suffix)).zero_()

        if cfg.rnn_layout == 0:            # stack_bidirectional_dynamic_rnn (chiron/rnn.py:64)
            m = torch.nn.LSTM(cfg.channels, H, num_layers=cfg.n_layers, batch_first=True, bidirectional=True)
            for l in range(cfg.n_layers):
                load_dir(m, l, "fw", "")
                load_dir(m, l, "bw", "_reverse")
            self.lstms.append(m.eval())
        else:                              # MultiRNNCell per direction + bidirectional_dynamic_rnn (chiron/rnn.py:140-145)
            for d in ("fw", "bw"):
                m = torch.nn.LSTM(cfg.channels, H, num_layers=cfg.n_layers, batch_first=True, bidirectional=False)
                for l in range(cfg.n_layers):
                    load_dir(m, l, d, "")
                self.lstms.append(m.eval())
        self.head_w = f(t["rnn_fnn_layer/weights"])
        self.head_b = f(t["rnn_fnn_layer/bias"])
        self.head_wc = f(t["rnn_fnn_layer/weights_class"])
        self.head_bc = f(t["rnn_fnn_layer/bias_class"])

    @staticmethod
    def _conv_same(x, w, stride):          # tf.nn.conv2d 'SAME' along time: x [B,C,T]
        T, k = x.shape[2], w.shape[2]
        t_out = -(-T // stride)
        pad = max((t_out - 1) * stride + k - T, 0)
        if pad:
            x = F.pad(x, (pad // 2, pad - pad // 2))
        return F.conv1d(x, w, stride=stride)

    @staticmethod
    def _bn(x, p):
        return x * p[0][None, :, None] + p[1][None, :, None]

    def cnn(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,L] -> [B,T,C] (chiron/cnn.py:234-262,334-389)."""
        net = x[:, None, :]
        for blk in self.blocks:
            b1 = self._conv_same(net, blk["w1"], blk["stride"])
            if blk["bn1"] is not None:
                b1 = self._bn(b1, blk["bn1"])
            a = torch.relu(self._bn(self._conv_same(net, blk["w2a"], 1), blk["bn2a"]))
            bb = torch.relu(self._bn(self._conv_same(a, blk["w2b"], blk["stride"]), blk["bn2b"]))
            c = self._bn(self._conv_same(bb, blk["w2c"], 1), blk["bn2c"])
            net = torch.relu(b1 + c)
        return net.transpose(1, 2).contiguous()

    def rnn(self, fea: torch.Tensor, lens: np.ndarray) -> torch.Tensor:
        """[B,T,C] -> lasth [B,T,2H]; dynamic_rnn sequence_length semantics through packed sequences (zero output and frozen
        state past the length; the backward direction reverses only the first len frames)."""
        B, T, _ = fea.shape
        lens_t = torch.as_tensor(np.asarray(lens), dtype=torch.int64)
        full = bool((lens_t == T).all())
        if self.cfg.rnn_layout == 0:
            if full:
                return self.lstms[0](fea)[0]
            return self._packed(self.lstms[0], fea, lens_t, T)
        outs = []
        for d, m in enumerate(self.lstms):
            x = fea
            if d == 1:
                x = self._reverse(fea, lens_t)
            y = m(x)[0] if full else self._packed(m, x, lens_t, T)
            outs.append(self._reverse(y, lens_t) if d == 1 else y)
        return torch.cat(outs, dim=2)

    @staticmethod
    def _packed(m, x, lens_t, T):
        keep = lens_t > 0
        out = torch.zeros(x.shape[0], T, m.hidden_size * (2 if m.bidirectional else 1))
        if keep.any():
            pk = torch.nn.utils.rnn.pack_padded_sequence(x[keep], lens_t[keep], batch_first=True, enforce_sorted=False)
            y, _ = torch.nn.utils.rnn.pad_packed_sequence(m(pk)[0], batch_first=True, total_length=T)
            out[keep] = y
        return out

    @staticmethod
    def _reverse(x, lens_t):               # array_ops.reverse_sequence
        B, T, _ = x.shape
        idx = torch.arange(T)[None, :].expand(B, T)
        src = torch.where(idx < lens_t[:, None], lens_t[:, None] - 1 - idx, idx)
        return torch.gather(x, 1, src[:, :, None].expand_as(x))

    def head(self, lasth: torch.Tensor) -> torch.Tensor:
        H = self.cfg.hidden
        h2 = lasth[:, :, :H] * self.head_w[0] + lasth[:, :, H:] * self.head_w[1] + self.head_b
        return h2 @ self.head_wc + self.head_bc

    @torch.no_grad()
    def inference(self, x: np.ndarray, lens_out: np.ndarray) -> np.ndarray:
        xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return self.head(self.rnn(self.cnn(xt), lens_out)).numpy()

    @torch.no_grad()
    def timed_pass(self, x: np.ndarray, lens_out: np.ndarray, decode) -> Tuple[Dict[str, float], np.ndarray, list]:
        """One pass with the per-stage split BASELINE.md section 3 asks for: seconds in the conv stack, the BiLSTM stack, the
        head + path_prob, and the CTC decode (``decode(logits, lens)`` -> list of label lists: the C oracle's decoder)."""
        xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        t0 = time.perf_counter()
        fea = self.cnn(xt)
        t1 = time.perf_counter()
        lasth = self.rnn(fea, lens_out)
        t2 = time.perf_counter()
        logits = self.head(lasth)
        top2 = torch.topk(logits, 2, dim=2).values
        prob = (top2[:, :, 0] - top2[:, :, 1]).mean(dim=1)
        t3 = time.perf_counter()
        lg = logits.numpy()
        paths = decode(lg, lens_out)
        t4 = time.perf_counter()
        return {"conv": t1 - t0, "lstm": t2 - t1, "head": t3 - t2, "decode": t4 - t3, "total": t4 - t0}, lg, paths
