"""Generate the assembly fixtures under tests/golden/assembly_ref/ by RUNNING THE REFERENCE'S OWN CODE in this container.

    python tools/gen_assembly_golden.py            # needs /root/reference (read-only); writes tests/golden/assembly_ref/*.json

chiron/utils/easy_assembler.py is pure Python (difflib + numpy); it is imported unmodified from /root/reference with two
shims -- a stub ``Bio.pairwise2`` module (only global_alignment_kernal uses it, which `chiron call` never selects) and
``np.lib.pad = np.pad`` (removed from numpy 2).  ``qs()`` lives in chiron/chiron_eval.py, which imports TensorFlow at module
level, so that one function is compiled from its source text (ast) without importing the module.  Nothing is copied into
the repository: only the inputs' names and the outputs are stored.

Third shim, an environment one: qs() calls ``np.argsort(consensus, axis=0)`` with the default kind.  On the numpy of the
reference's era (<= 1.16) a 4-element axis is insertion-sorted, i.e. STABLE, which decides whose quality sum is used where
the two highest base counts tie; numpy 2's SIMD argsort is not stable.  The reference's own golden result/read1.fastq
quality string is reproduced only by the stable order (tests/test_oracle_golden.py), so qs() is executed here against a
numpy proxy whose argsort defaults to kind="stable".  (Run with the plain numpy 2 argsort the fixtures differ from the
stable ones in 122 positions, every one of them an exact tie of the two highest counts.)

For each bundled golden segments file (tests/golden/DNA/segments/readN.fastq, FASTA-style records) and each assembly kernel
`chiron call` can select (simple / glue / stick, chiron_eval.py:138-150) the fixture holds the consensus of
simple_assembly(), and -- with per-window quality weights drawn from a seeded generator -- the consensus and the phred+33
string of simple_assembly_qs() + qs().  These pin the oracle's (and through it the CUDA path's) `simple` and `stick`
kernels and the quality-score arithmetic to the reference itself; the bundled result/*.fastq files only cover `glue`."""
import ast
import hashlib
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "assembly_ref")
CASES = [("simple", 0.5), ("simple", 390.0 / 400.0), ("glue", 390.0 / 400.0), ("stick", 1.0)]
MAX_SEGMENTS = 120           # bounds the fixture size and the O(len^2) difflib time of the `simple` kernel


def load_reference():
    bio = types.ModuleType("Bio")
    bio.pairwise2 = types.ModuleType("Bio.pairwise2")
    sys.modules.setdefault("Bio", bio)
    sys.modules.setdefault("Bio.pairwise2", bio.pairwise2)
    if not hasattr(np.lib, "pad"):
        np.lib.pad = np.pad
    sys.path.insert(0, REF)
    from chiron.utils import easy_assembler           # noqa: E402  (the reference, unmodified)
    src = open(os.path.join(REF, "chiron", "chiron_eval.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "qs")
    class _EraNumpy:                                   # numpy <= 1.16: argsort of a 4-element axis is stable (see above)
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def argsort(a, axis=-1, kind=None, order=None):
            return np.argsort(a, axis=axis, kind=kind or "stable", order=order)

    ns = {"np": _EraNumpy()}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "chiron_eval.qs", "exec"), ns)
    return easy_assembler, ns["qs"]


def read_segments(path):
    with open(path) as f:
        return [l.strip() for i, l in enumerate(f) if i % 2 == 1]


def overlapping_segments():
    """Segments that really overlap (what the `simple` kernel is for): read1 basecalled by the CPU oracle with the
    DNA_default weights at L=400, jump=100 (75 % overlap), greedy decoding, first 64 windows.  They are INPUTS of the
    fixture (stored in it); the assembly of them is the reference's."""
    sys.path.insert(0, ROOT)
    from chiron_b200.model import load_model
    from oracle import chiron_oracle as O
    cfg, t, _ = load_model("DNA_default")
    sig = O.read_signal_text(os.path.join(ROOT, "tests", "golden", "DNA", "raw", "read1.signal"))
    x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), 400, 100)
    x, lens = x[:64], lens[:64]
    logits = O.inference(x, lens, cfg, t)
    return [O.index2base(p) for p in O.ctc_greedy(logits, lens) if len(p)], 100.0 / 400.0


def main():
    ea, qs = load_reference()
    os.makedirs(OUT, exist_ok=True)
    index2base = np.array(list("ACGT"))
    segs, ratio = overlapping_segments()
    rng = np.random.default_rng(99)
    weights = rng.uniform(0.5, 12.0, size=len(segs)).astype(np.float32)
    cons_q, cons_qs = ea.simple_assembly_qs(segs, [np.array([w]) for w in weights], ratio, kernal="simple")
    seq = "".join(index2base[np.argmax(cons_q, axis=0)])
    assert seq == "".join(index2base[np.argmax(ea.simple_assembly(segs, ratio, kernal="simple"), axis=0)])
    with np.errstate(all="ignore"):
        quality = qs(cons_q, cons_qs)
    with open(os.path.join(OUT, "read1_overlap75.json"), "w") as f:
        json.dump({"generator": "tools/gen_assembly_golden.py", "reference": "chiron/utils/easy_assembler.py simple_assembly"
                   "(_qs) + chiron/chiron_eval.py qs(), run unmodified from /root/reference (qs: stable argsort, numpy<=1.16)",
                   "segments": segs, "segments_from": "CPU oracle, DNA_default, read1.signal, L=400, jump=100, greedy, 64 windows",
                   "n_segments": len(segs), "weights_seed": 99, "weights": [float(w) for w in weights],
                   "cases": [{"kernal": "simple", "jump_step_ratio": ratio, "consensus": seq, "quality": quality,
                              "consensus_sha256": hashlib.sha256(seq.encode()).hexdigest()}]}, f, indent=0)
    print("read1_overlap75: %d segments of ~%d bases -> %d bases" % (len(segs), np.mean([len(s) for s in segs]), len(seq)))
    for n in range(1, 6):
        segs = read_segments(os.path.join(ROOT, "tests", "golden", "DNA", "segments", "read%d.fastq" % n))[:MAX_SEGMENTS]
        rng = np.random.default_rng(100 + n)
        weights = rng.uniform(0.5, 12.0, size=len(segs)).astype(np.float32)
        cases = []
        for kernal, ratio in CASES:
            cons = ea.simple_assembly(segs, ratio, kernal=kernal)
            seq = "".join(index2base[np.argmax(cons, axis=0)])
            cons_q, cons_qs = ea.simple_assembly_qs(segs, [np.array([w]) for w in weights], ratio, kernal=kernal)    # rows of path_prob [B,1]
            seq_q = "".join(index2base[np.argmax(cons_q, axis=0)])
            with np.errstate(all="ignore"):
                quality = qs(cons_q, cons_qs)
            assert seq_q == seq
            cases.append({"kernal": kernal, "jump_step_ratio": ratio, "consensus": seq, "quality": quality,
                          "consensus_sha256": hashlib.sha256(seq.encode()).hexdigest()})
        with open(os.path.join(OUT, "read%d.json" % n), "w") as f:
            json.dump({"generator": "tools/gen_assembly_golden.py", "reference": "chiron/utils/easy_assembler.py simple_assembly"
                       "(_qs) + chiron/chiron_eval.py qs(), run unmodified from /root/reference (qs: stable argsort, numpy<=1.16)",
                       "segments_file": "tests/golden/DNA/segments/read%d.fastq" % n, "n_segments": len(segs),
                       "weights_seed": 100 + n, "weights": [float(w) for w in weights], "cases": cases}, f, indent=0)
        print("read%d: %d segments," % (n, len(segs)), ", ".join("%s@%.3f -> %d bases" % (c["kernal"], c["jump_step_ratio"],
                                                                                           len(c["consensus"])) for c in cases))


if __name__ == "__main__":
    main()
