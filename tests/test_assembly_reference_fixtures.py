"""The assembly path against fixtures produced by RUNNING THE REFERENCE'S OWN easy_assembler.simple_assembly(_qs) and
chiron_eval.qs() in the build container (tools/gen_assembly_golden.py; tests/golden/assembly_ref/*.json): all three kernels
`chiron call` can select (chiron_eval.py:138-150) on the five bundled golden segment files, plus really overlapping
segments for the `simple` kernel.  The bundled result/*.fastq files pin only the `glue` kernel; these pin `simple`, `stick`
and the quality-score arithmetic of every kernel to the reference itself.

Here (CPU): the oracle's restatement and the host-compiled displacement routines of the CUDA kernels reproduce every
fixture.  tests/test_gpu_zz_reference_fixtures.py (`-m gpu`): cb_assemble through the C ABI reproduces them too."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, read_fasta_records
from oracle import chiron_oracle as O

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "assembly_ref", "*.json")))
BASE_IDX = {"A": 0, "C": 1, "G": 2, "T": 3}


def _load(path):
    with open(path) as f:
        fx = json.load(f)
    if "segments" in fx:
        segs = fx["segments"]
    else:
        segs = read_fasta_records(os.path.join(os.path.dirname(GOLDEN), "..", fx["segments_file"]))[:fx["n_segments"]]
    assert len(segs) == fx["n_segments"]
    return fx, segs, np.asarray(fx["weights"], dtype=np.float32)


def test_fixture_set_is_complete():
    assert len(FIXTURES) == 6
    kernels = set()
    for path in FIXTURES:
        fx, segs, w = _load(path)
        assert len(w) == len(segs) and fx["generator"] == "tools/gen_assembly_golden.py"
        kernels |= {c["kernal"] for c in fx["cases"]}
    assert kernels == {"simple", "glue", "stick"}


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_oracle_reproduces_the_reference_assembler(path):
    fx, segs, w = _load(path)
    for case in fx["cases"]:
        cons, cons_qs, pos = O.simple_assembly_qs(segs, w, case["jump_step_ratio"], kernal=case["kernal"])
        assert O.index2base(np.argmax(cons, axis=0)) == case["consensus"], case["kernal"]
        assert O.qs_string(cons, cons_qs) == case["quality"], case["kernal"]
        cons2, _, pos2 = O.simple_assembly_qs(segs, None, case["jump_step_ratio"], kernal=case["kernal"])
        assert np.array_equal(cons2, cons) and np.array_equal(pos2, pos)


def test_device_displacement_routines_reproduce_the_reference_positions():
    """The sequential displacement routines the asm_disp kernel runs (host-compiled instantiation, cb_selftest_disp)
    rebuild every fixture's consensus when chained like simple_assembly does (easy_assembler.py:302-335)."""
    import ctypes
    from chiron_b200 import _lib
    lib = _lib.load()
    code = {"simple": _lib.ASM_SIMPLE, "glue": _lib.ASM_GLUE, "stick": _lib.ASM_STICK}
    for path in FIXTURES:
        fx, segs, _ = _load(path)
        enc = [np.array([BASE_IDX[c] for c in s], dtype=np.int8) for s in segs]
        for case in fx["cases"]:
            # the C routine takes jump and L; any pair with jump / L == jump_step_ratio is equivalent
            L = 400
            jump = int(round(case["jump_step_ratio"] * L))
            assert jump / L == case["jump_step_ratio"]
            pos, positions = 0, [0]
            for i in range(1, len(enc)):
                d = lib.cb_selftest_disp(enc[i].ctypes.data_as(ctypes.c_void_p), len(enc[i]),
                                         enc[i - 1].ctypes.data_as(ctypes.c_void_p), len(enc[i - 1]), code[case["kernal"]], jump, L)
                pos += d
                positions.append(pos)
            _, _, ref_pos = O.simple_assembly_qs(segs, None, case["jump_step_ratio"], kernal=case["kernal"])
            assert positions == ref_pos.tolist(), (os.path.basename(path), case["kernal"])
