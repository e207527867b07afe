"""Basecall and assemble the bundled read3 (319 windows, glue kernel): the workload of the ncu capture of the assembly kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from chiron_b200 import chiron_input
from chiron_b200.engine import Basecaller
from chiron_b200.model import load_model
cfg = load_model("DNA_default")[0]
data = chiron_input.read_data_for_eval("tests/golden/DNA/raw/read3.signal", 0, seg_length=400, step=390, sig_norm=cfg.sig_norm)
x, lens, _ = data.next_batch(data.reads_n, shuffle=False)
bc = Basecaller("DNA_default", 0)
bases, nb, prob, _ = bc.basecall_batch(x, lens, beam=0)
for jump, L in ((390, 400), (200, 400), (400, 400)):
    seq, qual, pos = bc.assemble(bases, nb, prob, jump, L)
    print(jump, L, len(seq))
