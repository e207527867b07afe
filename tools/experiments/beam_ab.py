"""A/B of the beam decoders on RNA config 3 (512 windows x 500 samples, beam 50) and DNA (400 x 400, beam 30)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from chiron_b200.engine import Basecaller
for model, B, L, W in (("RNA_default", 512, 500, 50), ("DNA_default", 400, 400, 30), ("DNA_default", 4096, 512, 30)):
    bc = Basecaller(model, 0, "tc")
    x = torch.randn(B, L, device="cuda") * 0.43 - 0.16
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    lo = bc.seq_len_out_device(lens, L)
    lg, _ = bc.forward_device(x, lo)
    torch.cuda.synchronize()
    for _ in range(2): bases, nb = bc.decode_device(lg, lo, beam=W)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): bases, nb = bc.decode_device(lg, lo, beam=W)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
    print(os.environ.get("CB_BEAM_SMEM", "1"), model, B, L, W, "beam decode ms %.2f" % ms, "checksum", int(bases.to(torch.int64).sum()), int(nb.sum()))
    bc.close()
