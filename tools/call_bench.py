"""End-to-end throughput of the `chiron call` path on FILES: folder of reads -> result/segments/meta files.

    python tools/call_bench.py [--reads N] [--fmt signal|fast5] [--precision tc] [--beam 0] [--stub]

Replicates the bundled reads (tests/golden) N times into a temporary folder, runs chiron_eval.run() on it with the
dna-pre preset (L=400, jump=390) and prints one JSON line: wall time, raw-signal Msamples/s and kbases/s from files to
fastq.  `--stub` swaps the GPU for a do-nothing Basecaller with the same interface, which measures the ceiling of the
HOST pipeline alone (reader threads, batching, collation, output formatting) -- runnable without a GPU."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from chiron_b200 import chiron_eval, chiron_input
from chiron_b200.model import load_model


class StubCaller:
    """Same surface as engine.Basecaller for evaluation(); returns ~20 bases per window without touching a GPU."""

    def __init__(self, model, bn_mode=0):
        self.cfg, _, _ = load_model(model)
        self._rng = np.random.default_rng(0)
        self.bn_mode = bn_mode
        self.precision = "stub"

    def out_len(self, L):
        return self.cfg.out_len(L)

    def basecall_submit(self, slot, x, seq_len, beam=0):
        return slot, x.shape[0], self.out_len(x.shape[1])

    def basecall_collect(self, ticket):
        _, B, T = ticket
        bases = self._rng.integers(0, 4, size=(B, T)).astype(np.int8)
        return bases, np.full(B, min(20, T), dtype=np.int32), np.full(B, 5.0, dtype=np.float32)

    def assemble(self, bases, n_bases, path_prob, jump, L, kernel=None, with_qs=True):
        n = int(n_bases.sum() * jump / L)
        return "A" * n, ("5" * n if with_qs else None), np.arange(len(n_bases), dtype=np.int32)

    def close(self):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=200)
    ap.add_argument("--fmt", default="signal", choices=["signal", "fast5"])
    ap.add_argument("--precision", default="tc")
    ap.add_argument("--beam", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--stub", action="store_true")
    a = ap.parse_args()
    if a.fmt == "signal":
        src = [os.path.join(ROOT, "tests", "golden", "DNA", "raw", n) for n in ("read1.signal", "read3.signal")]
    else:
        src = [os.path.join(ROOT, "tests", "golden", "fast5", "read1.fast5")]
    tmp = tempfile.mkdtemp(prefix="chiron_call_bench_")
    try:
        inp, out = os.path.join(tmp, "in"), os.path.join(tmp, "out")
        os.makedirs(inp)
        samples = 0
        per_src = [len(chiron_input.read_signal(p)) if p.endswith(".signal") else len(chiron_input.read_signal_fast5(p)) for p in src]
        for i in range(a.reads):
            j = i % len(src)
            shutil.copy(src[j], os.path.join(inp, "r%05d.%s" % (i, a.fmt)))
            samples += per_src[j]
        flags = types.SimpleNamespace(input=inp, output=out, model="DNA_default", start=None, batch_size=None, segment_len=None,
                                      jump=None, threads=a.threads or None, beam=a.beam, extension="fastq", concise=False,
                                      mode="dna", preset="dna-pre", precision=a.precision, recursive=False)
        flags = chiron_eval.apply_preset(flags)
        caller = StubCaller("DNA_default") if a.stub else None
        if not a.stub:                                     # warm the library (context, workspace) outside the timed region
            from chiron_b200.engine import Basecaller
            caller = Basecaller("DNA_default", device=0, precision=a.precision)
            caller.basecall_batch(np.zeros((4096, flags.segment_len), np.float32), np.full(4096, flags.segment_len, np.int32))
        os.makedirs(out, exist_ok=True)
        t0 = time.perf_counter()
        summary = chiron_eval.evaluation(flags, caller=caller)
        wall = time.perf_counter() - t0
        bases = sum(v["bases"] for v in summary.values())
        print(json.dumps({"what": "chiron call, files -> fastq (%s)" % ("host pipeline only, GPU stubbed" if a.stub else a.precision),
                          "reads": a.reads, "format": a.fmt, "samples": samples, "wall_s": round(wall, 3),
                          "Msamples_per_s": round(samples / wall / 1e6, 2), "kbases_per_s": round(bases / wall / 1e3, 1),
                          "reader_threads": a.threads or min(8, os.cpu_count() or 1), "host_cores": os.cpu_count(),
                          "beam": a.beam}))
        if caller is not None:
            caller.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
