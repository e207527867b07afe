"""Beam search on REAL logits (bootstrap windows of the bundled reads through the forward pass), for the next GPU session:
the default launcher (one pass, pool sized for the tail), the experimental two-pass search (CB_BEAM_RETRY=1) and the
thread-per-window fallback (CB_BEAM_SMEM=0).  One JSON line per (B, L, width): ms per decode, kernels launched per decode
(1 = the fast path held, 2 = retry pass or fallback ran, 3 = both) and whether the outputs are identical."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from bench import synthetic_windows
from chiron_b200.engine import Basecaller

MODES = {"default": {}, "retry": {"CB_BEAM_RETRY": "1"}, "fallback": {"CB_BEAM_SMEM": "0"}}

for B, L, W in ((4096, 512, 30), (4096, 400, 30), (512, 400, 30), (4096, 512, 50)):
    bc = Basecaller("DNA_default", 0, "tc")
    xs, lens_h = synthetic_windows(B, L, 4321)
    x, lens = torch.from_numpy(xs).cuda(), torch.from_numpy(lens_h).cuda()
    lo = bc.seq_len_out_device(lens, L)
    lg, _ = bc.forward_device(x, lo)
    torch.cuda.synchronize()
    res, outs = {}, {}
    for mode, env in MODES.items():
        for k in ("CB_BEAM_RETRY", "CB_BEAM_SMEM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        bases, nb = bc.decode_device(lg, lo, beam=W)            # warm-up (workspace, attributes)
        torch.cuda.synchronize()
        n0, t0 = bc.launches, time.perf_counter()
        reps = 1 if mode == "fallback" else 3
        for _ in range(reps):
            bases, nb = bc.decode_device(lg, lo, beam=W)
        torch.cuda.synchronize()
        res[mode] = {"ms": round((time.perf_counter() - t0) / reps * 1e3, 2), "launches_per_decode": (bc.launches - n0) / reps}
        outs[mode] = (bases.cpu(), nb.cpu())
    same = all(torch.equal(outs[m][0], outs["fallback"][0]) and torch.equal(outs[m][1], outs["fallback"][1]) for m in outs)
    print(json.dumps({"B": B, "L": L, "beam": W, "bases_per_window": round(float(outs["fallback"][1].float().mean()), 1),
                      "modes": res, "identical": same}), flush=True)
    bc.close()
