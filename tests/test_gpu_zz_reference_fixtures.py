"""GPU: cb_assemble (through the C ABI) against the fixtures produced by running the reference's own
easy_assembler.simple_assembly(_qs) + chiron_eval.qs() (tools/gen_assembly_golden.py, tests/golden/assembly_ref/): all three
assembly kernels on the five bundled golden segment files plus really overlapping segments.  The CPU side of the same
fixtures is tests/test_assembly_reference_fixtures.py."""
import os

import numpy as np
import pytest

from oracle import chiron_oracle as O
from test_assembly_reference_fixtures import BASE_IDX, FIXTURES, _load


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_cuda_assembly_reproduces_the_reference_assembler(path):
    from chiron_b200.engine import Basecaller
    fx, segs, w = _load(path)
    T = max(len(s) for s in segs)
    bases = np.zeros((len(segs), T), dtype=np.int8)
    n_bases = np.zeros(len(segs), dtype=np.int32)
    for i, s in enumerate(segs):
        bases[i, :len(s)] = [BASE_IDX[c] for c in s]
        n_bases[i] = len(s)
    bc = Basecaller("DNA_default", device=0)
    for case in fx["cases"]:
        L = 400
        jump = int(round(case["jump_step_ratio"] * L))
        seq, qual, pos = bc.assemble(bases, n_bases, w, jump, L, kernel=case["kernal"], with_qs=True)
        assert seq == case["consensus"], case["kernal"]
        # positions no window covers are undefined in both implementations (0/0 in qs()); compare the covered ones
        cons, _, _ = O.simple_assembly_qs(segs, w, case["jump_step_ratio"], kernal=case["kernal"])
        covered = cons.sum(axis=0) > 0
        assert [c for c, ok in zip(qual, covered) if ok] == [c for c, ok in zip(case["quality"], covered) if ok], case["kernal"]
    bc.close()
