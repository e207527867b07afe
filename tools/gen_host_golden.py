"""Generate tests/golden/host_ref/host_ref.json by RUNNING THE REFERENCE'S OWN pure-Python host functions.

    python tools/gen_host_golden.py                # needs /root/reference (read-only)

chiron/chiron_eval.py imports TensorFlow at module level and cannot be imported here, but write_output,
get_assembler_kernal, sparse2dense and index2base are plain Python: they are compiled one by one from the module's source
text (ast) and executed unmodified, with `time` replaced by a fixed clock so that the meta files are reproducible.
Nothing is copied into the repository; only arguments and outputs are stored.

The fixture pins rows a12, a13 and a16 of the path to the reference's own code:
  * write_output: result / segments / meta file contents for fastq, fasta, concise, rna mode and per-segment qualities;
  * get_assembler_kernal: the kernel chosen for a grid of (jump, segment_len);
  * sparse2dense + index2base on a SparseTensor-shaped greedy result with empty windows (the rows that disappear);
  * chiron_input.read_signal + read_data_for_eval + padding (rows a1-a2; compiled the same way from chiron/chiron_input.py,
    which imports h5py / statsmodels; FLAGS.sig_norm = None as at HEAD, DataSet replaced by a recorder of its arguments)
    on the bundled read1.signal for several (start, step, seg_length): window count, true lengths and a SHA-256 of the
    float32 window matrix."""
import ast
import collections
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EVAL = "/root/reference/chiron/chiron_eval.py"
REF_INPUT = "/root/reference/chiron/chiron_input.py"
OUT = os.path.join(ROOT, "tests", "golden", "host_ref", "host_ref.json")
FIXED_NOW = 1000.0


def load_functions(names):
    tree = ast.parse(open(REF_EVAL).read())
    fake_time = types.SimpleNamespace(time=lambda: FIXED_NOW)
    ns = {"np": np, "os": os, "time": fake_time}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), "chiron_eval." + node.name, "exec"), ns)
    return [ns[n] for n in names]


def load_input_functions():
    tree = ast.parse(open(REF_INPUT).read())
    ns = {"np": np, "MEAN": "mean", "MEDIAN": "median", "FLAGS": types.SimpleNamespace(sig_norm=None),
          "DataSet": lambda **kw: kw}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("read_signal", "padding", "read_data_for_eval"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "chiron_input." + node.name, "exec"), ns)
    return ns["read_data_for_eval"]


def main():
    write_output, get_assembler_kernal, sparse2dense, index2base = load_functions(
        ["write_output", "get_assembler_kernal", "sparse2dense", "index2base"])
    fixture = {"generator": "tools/gen_host_golden.py", "reference": "chiron/chiron_eval.py, functions run unmodified from "
               "/root/reference (time.time() fixed at %.1f)" % FIXED_NOW, "fixed_now": FIXED_NOW}

    # ---- write_output ------------------------------------------------------------------------------------------------------
    segs = ["ACGTTGCA", "GGCAT", "T", "CCATGCATGCAAGT"]
    consensus = "ACGTTGCAGGCATTCCATGCATGCAAGT"
    qual = "".join(chr(33 + (7 * i) % 40) for i in range(len(consensus)))
    seg_q = ["".join(chr(40 + (3 * i + j) % 30) for j in range(len(s))) for i, s in enumerate(segs)]
    times = [FIXED_NOW - 12.5, 0.25, 3.5, 4.75]                  # start_time, reading, basecall, assembly (cumulative)
    cases = []
    for name, kw in [("fastq_with_quality", dict(suffix="fastq", q_score=qual)),
                     ("fastq_without_quality", dict(suffix="fastq")),
                     ("fasta", dict(suffix="fasta")),
                     ("concise_fastq", dict(suffix="fastq", q_score=qual, concise=True)),
                     ("rna_mode", dict(suffix="fastq", q_score=qual, mode="rna")),
                     ("segment_qualities", dict(suffix="fastq", q_score=qual, seg_q_score=seg_q)),
                     ("no_segments", dict(suffix="fastq", q_score="", segments=[], consensus=""))]:
        with tempfile.TemporaryDirectory() as tmp:
            for sub in ("result", "segments", "meta"):
                os.makedirs(os.path.join(tmp, sub))
            gs = types.SimpleNamespace(output=tmp, mode=kw.pop("mode", "dna"), batch_size=400, segment_len=400, jump=390, start=0,
                                       input="/data/in/", model="DNA_default")
            s_arg, c_arg = kw.pop("segments", segs), kw.pop("consensus", consensus)
            write_output(s_arg, c_arg, list(times), "read_x", global_setting=gs, **kw)
            files = {}
            for dirpath, _, fns in os.walk(tmp):
                for fn in fns:
                    with open(os.path.join(dirpath, fn), newline="") as f:
                        files[os.path.relpath(os.path.join(dirpath, fn), tmp)] = f.read()
            cases.append({"name": name, "segments": s_arg, "consensus": c_arg, "time_list": times, "file_pre": "read_x",
                          "mode": gs.mode, "kwargs": kw, "files": files})
    fixture["write_output"] = cases

    # ---- get_assembler_kernal --------------------------------------------------------------------------------------------------
    grid = []
    for L in (300, 400, 500, 512, 2000):
        for jump in sorted({1, L // 2, int(0.9 * L) - 1, int(0.9 * L), int(0.9 * L) + 1, L - 10, L - 1, L, L + 1, 2 * L, 390, 440, 1900}):
            grid.append([jump, L, get_assembler_kernal(jump, L)])
    fixture["get_assembler_kernal"] = grid

    # ---- sparse2dense + index2base ---------------------------------------------------------------------------------------------
    rng = np.random.default_rng(42)
    B, T = 23, 17
    n_bases = rng.integers(0, T + 1, size=B)
    n_bases[[0, 5, 22]] = 0                                       # windows that decode to nothing disappear
    bases = rng.integers(0, 4, size=(B, T))
    idx = np.array([[b, t] for b in range(B) for t in range(n_bases[b])], dtype=np.int64).reshape(-1, 2)
    vals = np.array([bases[b, t] for b in range(B) for t in range(n_bases[b])], dtype=np.int64)
    Sparse = collections.namedtuple("SparseTensorValue", ["indices", "values", "dense_shape"])
    predict_read, uniq = sparse2dense(([Sparse(idx, vals, np.array([B, T]))], np.zeros((B, 1), np.float32)))
    fixture["sparse2dense"] = {"bases": bases.tolist(), "n_bases": n_bases.tolist(),
                               "reads": [index2base(r) for r in predict_read[0]], "uniq": [int(u) for u in uniq[0]]}
    # ---- read_signal + read_data_for_eval + padding ---------------------------------------------------------------------------
    import hashlib
    read_data_for_eval = load_input_functions()
    sig_path = os.path.join(ROOT, "tests", "golden", "DNA", "raw", "read1.signal")
    windows = []
    for start, step, seg in ((0, 390, 400), (0, 290, 300), (7, 100, 300), (0, 512, 512), (62000, 5, 100000), (0, 1900, 2000)):
        ds = read_data_for_eval(sig_path, start_index=start, step=step, seg_length=seg)
        event = np.asarray(ds["event"], dtype=np.float32)
        windows.append({"start_index": start, "step": step, "seg_length": seg, "n_windows": int(event.shape[0]),
                        "event_length": [int(v) for v in ds["event_length"]],
                        "event_sha256": hashlib.sha256(np.ascontiguousarray(event).tobytes()).hexdigest(),
                        "first_window_head": [float(v) for v in event[0, :8]], "last_window_sum": float(event[-1].sum())})
    fixture["read_data_for_eval"] = {"file": "tests/golden/DNA/raw/read1.signal", "sig_norm": None, "cases": windows}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(fixture, f, indent=0)
    print("wrote %s: %d write_output cases, %d kernel choices, %d dense reads" % (
        OUT, len(cases), len(grid), len(fixture["sparse2dense"]["reads"])))


if __name__ == "__main__":
    main()
