"""Beam search (width 30) on real logits, default launcher: median ms of 5 decodes at 4096 x 512 and 4096 x 400."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench import synthetic_windows
from chiron_b200.engine import Basecaller
bc = Basecaller("DNA_default", 0, "tc")
for L in (512, 400):
    xs, lens_h = synthetic_windows(4096, L, 4321)
    x, lens = torch.from_numpy(xs).cuda(), torch.from_numpy(lens_h).cuda()
    lo = bc.seq_len_out_device(lens, L)
    lg, _ = bc.forward_device(x, lo)
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bases, nb = bc.decode_device(lg, lo, beam=30); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(L, "ms %.2f" % sorted(ts[1:])[2], "checksum", int(bases.to(torch.int64).sum()), int(nb.sum()))
