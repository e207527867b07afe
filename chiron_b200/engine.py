"""Python face of the C ABI: one ``Basecaller`` per GPU.

Replaces ``build_eval_graph`` / ``net`` (chiron/chiron_eval.py:244-376): the TF session + queues become a handle on
libchiron_b200.so.  Host (numpy) calls go through ``cb_basecall_host`` / ``cb_assemble_host``; device calls take torch
CUDA tensors (torch is plumbing for device memory and streams only) and pass raw pointers."""
from __future__ import annotations

import ctypes
import sys
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .model import load_model

BASES = "ACGT"
_BASE_LUT = np.frombuffer(BASES.encode("ascii"), dtype=np.uint8)


def index2base(read: Sequence[int]) -> str:
    """chiron/chiron_eval.py:100-113 (one table lookup over the whole read instead of a Python loop per base)."""
    idx = np.asarray(read, dtype=np.int64)
    if idx.size and (idx.min() < 0 or idx.max() > 3):
        raise IndexError("base index outside 0..3")
    return _BASE_LUT[idx].tobytes().decode("ascii")


def windows2bases(bases: np.ndarray, n_bases: np.ndarray) -> List[str]:
    """index2base of every non-empty row of a dense [n, T] decode result (sparse2dense drops the empty rows,
    chiron/chiron_eval.py:56-66)."""
    chars = _BASE_LUT[np.asarray(bases, dtype=np.int64) & 3]
    return [chars[i, :n_bases[i]].tobytes().decode("ascii") for i in np.nonzero(np.asarray(n_bases) > 0)[0]]


def get_assembler_kernal(jump: int, segment_len: int) -> str:
    """chiron/chiron_eval.py:138-150 (same name, same thresholds)."""
    assembler = "simple"
    if jump > 0.9 * segment_len:
        assembler = "glue"
    if jump >= segment_len:
        assembler = "stick"
    return assembler


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def format_segments(file_pre: str, bases: np.ndarray, n_bases: np.ndarray) -> bytes:
    """The text of segments/<file_pre>.<ext> for a dense [n, T] decode result: ``">{}{}\\n{}\\n".format(file_pre, idx,
    read)`` per non-empty window (write_output, chiron/chiron_eval.py:211-214), formatted natively without the GIL."""
    bases = np.ascontiguousarray(bases, dtype=np.int8)
    n_bases = np.ascontiguousarray(n_bases, dtype=np.int32)
    n, T = bases.shape
    name = file_pre.encode("utf-8")
    kept = int((n_bases > 0).sum())
    cap = kept * (len(name) + 13) + int(np.clip(n_bases, 0, T).sum()) + 1
    buf = ctypes.create_string_buffer(cap)
    got = _lib.load().cb_host_format_segments(name, _ptr(bases), _ptr(n_bases), n, T, buf, cap)
    if got < 0:
        _lib.check(int(got), "cb_host_format_segments")
    return buf.raw[:got]


class Basecaller:
    def __init__(self, model: str = "DNA_default", device: int = 0, precision: str = "auto",
                 bn_mode: Optional[str] = None):
        """``precision``: "tc" = the tcgen05 tensor-core kernels (the production mode), "fp32" = the FFMA kernels,
        "auto" (default) = "tc" whenever the model's topology is one the tensor-core kernels cover (the shipped
        DNA_default / RNA_default, residual stacks of any depth, width and stride, the stem models RNA_model2/3),
        "fp32" otherwise (GRU cells, hidden != 100, batch-statistics BatchNorm) -- said once on stderr, never silently.
        ``bn_mode``: None = what the model's blob header says; "population" / "batch" override it
        (cb_set_bn_mode; "batch" = HEAD's simple_global_bn, chiron/cnn.py:166-188, fp32 precision only)."""
        self.lib = _lib.load()
        self.cfg, _, blob = load_model(model)
        self.device = int(device)
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        if precision == "auto":
            wants_batch_bn = bn_mode == "batch" or (bn_mode is None and getattr(self.cfg, "bn_mode", 0) == _lib.BN_BATCH)
            rc = _lib.CB_ERR_ARG if wants_batch_bn else self.lib.cb_create(
                ctypes.cast(buf, ctypes.c_void_p), len(blob), self.device, _lib.PREC_TC_SPLIT, ctypes.byref(h))
            if rc == _lib.CB_OK:
                precision = "tc"
            elif rc == _lib.CB_ERR_ARG:
                why = "batch-statistics BatchNorm" if wants_batch_bn else self.lib.cb_last_error().decode("utf-8", "replace")
                sys.stderr.write("chiron_b200: model %s runs on the fp32 kernels (%s)\n" % (model, why))
                precision = "fp32"
            else:
                _lib.check(rc, "cb_create")
        self.precision = precision
        if not h:
            _lib.check(self.lib.cb_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), self.device,
                                          _lib.PRECISIONS[precision], ctypes.byref(h)), "cb_create")
        self.h = h
        if bn_mode is not None:
            rc = self.lib.cb_set_bn_mode(self.h, _lib.BN_MODES[bn_mode])
            if rc != _lib.CB_OK:
                msg = self.lib.cb_last_error().decode("utf-8", "replace")
                self.close()
                raise _lib.ChironB200Error("cb_set_bn_mode failed (%d): %s" % (rc, msg))
        self.bn_mode = self.lib.cb_bn_mode(self.h)
        self.n_class = self.lib.cb_n_class(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- facts -------------------------------------------------------------------------------------------------
    def out_len(self, L: int) -> int:
        return self.lib.cb_out_len(self.h, int(L))

    @property
    def launches(self) -> int:
        return int(self.lib.cb_launch_count(self.h))

    def reserve_sms(self, n: int):
        """Leave ``n`` SMs out of the persistent contraction grids (cb_reserve_sms): for pipelines that run small kernels
        of another stream next to the forward pass, like evaluation()'s per-read assembly."""
        _lib.check(self.lib.cb_reserve_sms(self.h, int(n)), "cb_reserve_sms")

    def enable_timing(self, on: bool = True):
        self.lib.cb_enable_timing(self.h, int(on))

    def last_forward_ms(self) -> List[float]:
        arr = (ctypes.c_float * 4)()
        n = self.lib.cb_last_forward_ms(self.h, arr, 4)
        return [float(arr[i]) for i in range(n)]

    def last_forward_profile(self):
        """{category: (summed ms, launches)} of the last forward (timing enabled, stream synchronised)."""
        ms = (ctypes.c_float * 4)()
        cnt = (ctypes.c_int * 4)()
        n = self.lib.cb_last_forward_profile(self.h, ms, cnt, 4)
        names = ["conv", "lstm_in", "lstm_rec", "head"]
        return {names[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}

    # ---- host (numpy) path: what `evaluation()` uses -----------------------------------------------------------
    def basecall_batch(self, x: np.ndarray, seq_len: np.ndarray, beam: int = 0, want_logits: bool = False):
        """x [B,L] float32 windows, seq_len [B] int32 true lengths.  Returns (bases[B,T] int8, n_bases[B],
        path_prob[B], logits[B,T,C] or None)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.int32)
        B, L = x.shape
        T = self.out_len(L)
        bases = np.zeros((B, T), dtype=np.int8)
        n_bases = np.zeros(B, dtype=np.int32)
        prob = np.zeros(B, dtype=np.float32)
        logits = np.zeros((B, T, self.n_class), dtype=np.float32) if want_logits else None
        _lib.check(self.lib.cb_basecall_host(self.h, _ptr(x), _ptr(seq_len), B, L, int(beam), _ptr(bases),
                                             _ptr(n_bases), _ptr(prob), _ptr(logits)), "cb_basecall_host")
        return bases, n_bases, prob, logits

    def basecall_submit(self, slot: int, x: np.ndarray, seq_len: np.ndarray, beam: int = 0):
        """Asynchronous two-slot form of ``basecall_batch`` (cb_basecall_submit): returns as soon as the batch sits in
        the slot's pinned staging buffer and its copies/kernels are enqueued.  Returns the ticket ``collect`` needs."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.int32)
        B, L = x.shape
        _lib.check(self.lib.cb_basecall_submit(self.h, int(slot), _ptr(x), _ptr(seq_len), B, L, int(beam)),
                   "cb_basecall_submit")
        return int(slot), B, self.out_len(L)

    def basecall_collect(self, ticket):
        """Wait for a submitted batch: (bases[B,T] int8, n_bases[B], path_prob[B])."""
        slot, B, T = ticket
        bases = np.zeros((B, T), dtype=np.int8)
        n_bases = np.zeros(B, dtype=np.int32)
        prob = np.zeros(B, dtype=np.float32)
        _lib.check(self.lib.cb_basecall_collect(self.h, slot, _ptr(bases), _ptr(n_bases), _ptr(prob)),
                   "cb_basecall_collect")
        return bases, n_bases, prob

    def assemble(self, bases: np.ndarray, n_bases: np.ndarray, path_prob: Optional[np.ndarray], jump: int, L: int,
                 kernel: Optional[str] = None, with_qs: bool = True) -> Tuple[str, Optional[str], np.ndarray]:
        """simple_assembly(_qs) + argmax + qs() for one read; windows in true order.  Returns (sequence, quality
        string or None, pos[n_windows])."""
        bases = np.ascontiguousarray(bases, dtype=np.int8)
        n_bases = np.ascontiguousarray(n_bases, dtype=np.int32)
        n, T = bases.shape
        kernel = kernel or get_assembler_kernal(jump, L)
        max_len = int(n_bases.sum()) + 1
        cons = np.zeros(max_len, dtype=np.int8)
        qual = np.zeros(max_len, dtype=np.uint8) if with_qs else None
        pos = np.zeros(max(n, 1), dtype=np.int32)
        out_len = np.zeros(1, dtype=np.int32)
        pp = np.ascontiguousarray(path_prob, dtype=np.float32) if path_prob is not None else None
        if with_qs and pp is None:
            raise ValueError("quality scores need path_prob")
        _lib.check(self.lib.cb_assemble_host(self.h, _ptr(bases), _ptr(n_bases), _ptr(pp), n, T, int(jump), int(L),
                                             _lib.ASM_KERNELS[kernel], _ptr(cons), _ptr(qual), _ptr(pos),
                                             _ptr(out_len), max_len), "cb_assemble_host")
        ln = int(out_len[0])
        seq = index2base(cons[:ln])
        q = bytes(qual[:ln]).decode("latin-1") if with_qs else None
        return seq, q, pos[:n]

    # ---- device (torch tensor) path: bench / multi-stream pipelines ----------------------------------------------
    def forward_device(self, x, seq_len_out, logits=None, path_prob=None, stream=None):
        """x [B,L] float32 CUDA tensor, seq_len_out [B] int32 CUDA tensor (already divided by ratio)."""
        import torch
        B, L = x.shape
        T = self.out_len(L)
        if logits is None:
            logits = torch.empty((B, T, self.n_class), dtype=torch.float32, device=x.device)
        if path_prob is None:
            path_prob = torch.empty((B,), dtype=torch.float32, device=x.device)
        s = stream if stream is not None else torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self.lib.cb_forward(self.h, x.data_ptr(), seq_len_out.data_ptr(), B, L, logits.data_ptr(),
                                       path_prob.data_ptr(), ctypes.c_void_p(s)), "cb_forward")
        return logits, path_prob

    def seq_len_out_device(self, seq_len_in, L: int, out=None, stream=None):
        import torch
        B = seq_len_in.shape[0]
        if out is None:
            out = torch.empty((B,), dtype=torch.int32, device=seq_len_in.device)
        s = stream if stream is not None else torch.cuda.current_stream(seq_len_in.device).cuda_stream
        _lib.check(self.lib.cb_seq_len_out(self.h, seq_len_in.data_ptr(), B, int(L), out.data_ptr(),
                                           ctypes.c_void_p(s)), "cb_seq_len_out")
        return out

    def decode_device(self, logits, seq_len_out, beam: int = 0, bases=None, n_bases=None, stream=None, log_prob=None):
        """``log_prob``: a float32 [B] device tensor that receives the beam search's top-path log probability
        (cb_decode_beam_scored; beam > 0 only)."""
        import torch
        B, T, _ = logits.shape
        if bases is None:
            bases = torch.empty((B, T), dtype=torch.int8, device=logits.device)
        if n_bases is None:
            n_bases = torch.empty((B,), dtype=torch.int32, device=logits.device)
        s = stream if stream is not None else torch.cuda.current_stream(logits.device).cuda_stream
        if beam == 0:
            _lib.check(self.lib.cb_decode_greedy(self.h, logits.data_ptr(), seq_len_out.data_ptr(), B, T,
                                                 bases.data_ptr(), n_bases.data_ptr(), ctypes.c_void_p(s)),
                       "cb_decode_greedy")
        elif log_prob is not None:
            _lib.check(self.lib.cb_decode_beam_scored(self.h, logits.data_ptr(), seq_len_out.data_ptr(), B, T, int(beam),
                                                      bases.data_ptr(), n_bases.data_ptr(), log_prob.data_ptr(),
                                                      ctypes.c_void_p(s)), "cb_decode_beam_scored")
        else:
            _lib.check(self.lib.cb_decode_beam(self.h, logits.data_ptr(), seq_len_out.data_ptr(), B, T, int(beam),
                                               bases.data_ptr(), n_bases.data_ptr(), ctypes.c_void_p(s)),
                       "cb_decode_beam")
        return bases, n_bases

    def predict(self, x: np.ndarray, seq_len: np.ndarray, beam_width: int = 30) -> dict:
        """The reference's serving signature as one call (chiron/export_test.py:24-40,103-113; what
        chiron/chiron_client.py:207-228 sends and reads back): ``{x [N,L] f32, seq_len [N] i32}`` ->
        ``{indices, values, dense_shape, logits, prob_logits, log_prob}``.

        ``indices [n,2] / values [n] / dense_shape [2]`` (int64) are the SparseTensor ``predict[0]`` of
        tf.nn.ctc_beam_search_decoder(merge_repeated=False) in row-major order, ``logits [N,T,C]``, ``prob_logits [N]`` =
        path_prob(logits), ``log_prob [N,1]`` = the decoder's score of the top path.  ``seq_len`` is divided by the model's
        stride ratio with tf.round like output_list() does (cb_seq_len_out).  Runs forward + beam search on the device;
        host<->device copies go through torch (plumbing)."""
        import torch
        if beam_width < 1:
            raise ValueError("the serving signature decodes with a beam search: beam_width >= 1")
        x = np.ascontiguousarray(x, dtype=np.float32)
        seq_len = np.ascontiguousarray(seq_len, dtype=np.int32).reshape(-1)
        if x.ndim != 2 or seq_len.shape[0] != x.shape[0]:
            raise ValueError("x must be [N, L] and seq_len [N]")
        N, L = x.shape
        dev = torch.device("cuda", self.device)
        with torch.cuda.device(dev):
            dx = torch.from_numpy(x).to(dev)
            dlen = self.seq_len_out_device(torch.from_numpy(seq_len).to(dev), L)
            logits, prob = self.forward_device(dx, dlen)
            score = torch.zeros((N,), dtype=torch.float32, device=dev)
            bases, n_bases = self.decode_device(logits, dlen, beam=int(beam_width), log_prob=score)
            self.check_status()
            bases, n_bases = bases.cpu().numpy(), n_bases.cpu().numpy()
            out_logits, prob, score = logits.cpu().numpy(), prob.cpu().numpy(), score.cpu().numpy()
        T = bases.shape[1]
        keep = np.arange(T)[None, :] < n_bases[:, None]
        rows, cols = np.nonzero(keep)                      # row-major: the order TF emits the sparse entries in
        return {"indices": np.stack([rows, cols], axis=1).astype(np.int64),
                "values": bases[rows, cols].astype(np.int64),
                "dense_shape": np.array([N, int(n_bases.max()) if N else 0], dtype=np.int64),
                "logits": out_logits, "prob_logits": prob, "log_prob": score.reshape(N, 1)}

    def check_status(self, stream=None):
        """Synchronise the stream and raise the deferred device-side errors of the asynchronous calls (cb_check_status):
        an activation outside the fp16 range of the tensor-core path, a beam search that ran out of fallback workspaces."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        _lib.check(self.lib.cb_check_status(self.h, ctypes.c_void_p(s)), "cb_check_status")

    def debug_fetch(self, what: int, n_floats: int) -> np.ndarray:
        out = np.zeros(n_floats, dtype=np.float32)
        n = self.lib.cb_debug_fetch(self.h, int(what), _ptr(out), n_floats)
        if n < 0:
            _lib.check(int(n), "cb_debug_fetch")
        return out[:n]
