"""Host-side rows a12, a13 and a16 against fixtures produced by RUNNING THE REFERENCE'S OWN chiron_eval.write_output,
get_assembler_kernal, sparse2dense and index2base (tools/gen_host_golden.py -> tests/golden/host_ref/host_ref.json)."""
import json
import os
import types

import numpy as np
import pytest

from conftest import GOLDEN
from chiron_b200 import chiron_eval, engine

with open(os.path.join(GOLDEN, "host_ref", "host_ref.json")) as _f:
    FX = json.load(_f)


def _run_write_output(case, tmp_path, monkeypatch, segments):
    for sub in ("result", "segments", "meta"):
        os.makedirs(os.path.join(str(tmp_path), sub), exist_ok=True)
    gs = types.SimpleNamespace(output=str(tmp_path), mode=case["mode"], batch_size=400, segment_len=400, jump=390, start=0,
                               input="/data/in/", model="DNA_default")
    monkeypatch.setattr(chiron_eval.time, "time", lambda: FX["fixed_now"])
    chiron_eval.write_output(segments, case["consensus"], list(case["time_list"]), case["file_pre"], global_setting=gs,
                             **case["kwargs"])
    files = {}
    for dirpath, _, fns in os.walk(str(tmp_path)):
        for fn in fns:
            with open(os.path.join(dirpath, fn), newline="") as f:
                files[os.path.relpath(os.path.join(dirpath, fn), str(tmp_path))] = f.read()
    return files


@pytest.mark.parametrize("case", FX["write_output"], ids=[c["name"] for c in FX["write_output"]])
def test_write_output_files_are_byte_identical_to_the_reference(case, tmp_path, monkeypatch):
    assert _run_write_output(case, tmp_path, monkeypatch, case["segments"]) == case["files"]


@pytest.mark.parametrize("case", [c for c in FX["write_output"] if "seg_q_score" not in c["kwargs"]],
                         ids=[c["name"] for c in FX["write_output"] if "seg_q_score" not in c["kwargs"]])
def test_native_segment_records_give_the_same_files(case, tmp_path, monkeypatch):
    """What evaluation() really passes: the records formatted by cb_host_format_segments from the dense decode result."""
    segs = case["segments"]
    T = max([len(s) for s in segs] + [1])
    bases = np.zeros((len(segs) + 2, T), dtype=np.int8)
    n_bases = np.zeros(len(segs) + 2, dtype=np.int32)          # two empty windows mixed in: they must leave no record
    rows = [i for i in range(len(segs) + 2) if i not in (1, len(segs) + 1)]
    for r, s in zip(rows, segs):
        bases[r, :len(s)] = ["ACGT".index(c) for c in s]
        n_bases[r] = len(s)
    records = engine.format_segments(case["file_pre"], bases, n_bases)
    assert _run_write_output(case, tmp_path, monkeypatch, records) == case["files"]


def test_kernel_choice_matches_the_reference():
    for jump, L, kernal in FX["get_assembler_kernal"]:
        assert engine.get_assembler_kernal(jump, L) == kernal, (jump, L)


def test_dense_rows_to_reads_matches_sparse2dense_and_index2base():
    fx = FX["sparse2dense"]
    bases = np.asarray(fx["bases"], dtype=np.int8)
    n_bases = np.asarray(fx["n_bases"], dtype=np.int32)
    assert engine.windows2bases(bases, n_bases) == fx["reads"]
    assert np.nonzero(n_bases > 0)[0].tolist() == fx["uniq"]          # the rows sparse2dense keeps
    assert [engine.index2base(bases[b, :n_bases[b]]) for b in fx["uniq"]] == fx["reads"]


def test_signal_windows_match_the_reference_reader():
    """chiron_input.read_signal + read_data_for_eval + padding run from the reference (sig_norm None as at HEAD) against
    the native parse + window path, for several (start, step, seg_length) incl. windows longer than the read."""
    import hashlib
    from chiron_b200 import chiron_input
    from chiron_b200.model import NORM_NONE
    fx = FX["read_data_for_eval"]
    path = os.path.join(os.path.dirname(GOLDEN), "..", fx["file"])
    for case in fx["cases"]:
        ds = chiron_input.read_data_for_eval(path, case["start_index"], step=case["step"], seg_length=case["seg_length"],
                                             sig_norm=NORM_NONE)
        assert ds.reads_n == case["n_windows"]
        assert ds.event_length.tolist() == case["event_length"]
        assert ds.event.dtype == np.float32 and ds.event.shape == (case["n_windows"], case["seg_length"])
        assert hashlib.sha256(np.ascontiguousarray(ds.event).tobytes()).hexdigest() == case["event_sha256"]
        assert ds.event[0, :8].tolist() == case["first_window_head"]
