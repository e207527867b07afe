"""How many trie nodes does the beam search need on REAL logits?  (The evidence behind cb_seq_kernels.cuh: beam_small_pool.)

    python tools/beam_pool_survey.py [--width 30] [--logits path/to/logits.npy]

Runs the library's own sequential beam search (cb_selftest_beam, the host-compiled instantiation of the routine the kernels
run) on every window with increasing pool sizes and prints, per tier, how many windows need it.  Inputs: the reference's
chiron/utils/logits_sample.npy when /root/reference is present (1100 windows, T=300), else the 24-window slice under
tests/golden/logits; plus --oracle N windows of the bundled read3 basecalled by the CPU oracle at L=512 (slow)."""
import argparse
import collections
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from chiron_b200 import _lib


def survey(lg, lens, W, tiers):
    lib = _lib.load()
    B, T, C = lg.shape
    out = np.zeros(T, np.int8)
    cnt = collections.Counter()
    for b in range(B):
        row = np.ascontiguousarray(lg[b])
        for m in tiers:
            n = lib.cb_selftest_beam(row.ctypes.data_as(ctypes.c_void_p), int(lens[b]), C, W, max(m * W, 2 * W + 2),
                                     out.ctypes.data_as(ctypes.c_void_p))
            if n >= 0:
                cnt[m] += 1
                break
        else:
            cnt["more"] += 1
    return {str(k): cnt[k] for k in list(tiers) + ["more"] if cnt[k]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=30)
    ap.add_argument("--logits", default=None)
    ap.add_argument("--oracle", type=int, default=0, help="also survey N oracle-basecalled 512-sample windows of read3")
    a = ap.parse_args()
    tiers = (3, 6, 8, 12, 16, 24, 32, 48)
    path = a.logits or "/root/reference/chiron/utils/logits_sample.npy"
    if not os.path.exists(path):
        path = os.path.join(ROOT, "tests", "golden", "logits", "logits_sample_24.npy")
    lg = np.load(path)
    res = {"beam_width": a.width, "pool_tiers_x_width": list(tiers),
           os.path.basename(path): {"windows": int(lg.shape[0]), "T": int(lg.shape[1]),
                                    "windows_by_smallest_sufficient_tier": survey(lg, np.full(lg.shape[0], lg.shape[1]), a.width, tiers)}}
    if a.oracle:
        from chiron_b200.model import load_model
        from oracle import chiron_oracle as O
        cfg, t, _ = load_model("DNA_default")
        sig = O.read_signal_text(os.path.join(ROOT, "tests", "golden", "DNA", "raw", "read3.signal"))
        x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), 512, 512)
        x, lens = x[:a.oracle], lens[:a.oracle]
        res["read3_L512_oracle"] = {"windows": int(len(x)), "T": 512,
                                    "windows_by_smallest_sufficient_tier": survey(O.inference(x, lens, cfg, t), lens, a.width, tiers)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
