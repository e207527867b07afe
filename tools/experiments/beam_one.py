"""One beam-search decode (width 30) of a 4096 x 512 batch of real logits: the workload of the ncu capture of the beam kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench import synthetic_windows
from chiron_b200.engine import Basecaller
bc = Basecaller("DNA_default", 0, "tc")
xs, lens_h = synthetic_windows(4096, 512, 4321)
x, lens = torch.from_numpy(xs).cuda(), torch.from_numpy(lens_h).cuda()
lo = bc.seq_len_out_device(lens, 512)
lg, _ = bc.forward_device(x, lo)
bases, nb = bc.decode_device(lg, lo, beam=30)
torch.cuda.synchronize()
bc.check_status()
print("beam 30:", float(nb.float().mean()), "bases per window")
