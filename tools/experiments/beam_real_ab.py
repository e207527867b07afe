"""Beam search on REAL logits (bootstrap windows of the bundled reads through the forward pass): the three-pass launcher
with first-pass pools of 8W / 12W / the default (16W grown into the occupancy slack) / 24W nodes, and the thread-per-window
kernel over global workspaces (CB_BEAM_SMEM=0).  One JSON line per (B, L, width): ms per decode (CUDA events on the launching
stream), how many windows the first pass left marked, and whether all outputs are identical."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

from bench import synthetic_windows
from chiron_b200.engine import Basecaller

CASES = ((4096, 512, 30), (4096, 400, 30), (512, 400, 30), (4096, 512, 50), (1024, 512, 30))
for B, L, W in CASES:
    modes = {"pool_8W": {"CB_BEAM_POOL": str(8 * W)}, "pool_12W": {"CB_BEAM_POOL": str(12 * W)}, "default": {},
             "pool_24W": {"CB_BEAM_POOL": str(24 * W)}, "global_kernel": {"CB_BEAM_SMEM": "0"}}
    bc = Basecaller("DNA_default", 0, "tc")
    xs, lens_h = synthetic_windows(B, L, 4321)
    x, lens = torch.from_numpy(xs).cuda(), torch.from_numpy(lens_h).cuda()
    lo = bc.seq_len_out_device(lens, L)
    lg, _ = bc.forward_device(x, lo)
    torch.cuda.synchronize()
    res, outs = {}, {}
    for mode, env in modes.items():
        for k in ("CB_BEAM_POOL", "CB_BEAM_SMEM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        bases, nb = bc.decode_device(lg, lo, beam=W)            # warm-up (workspace, attributes)
        torch.cuda.synchronize()
        reps = 1 if mode == "global_kernel" else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = bc.launches
        e0.record()
        for _ in range(reps):
            bases, nb = bc.decode_device(lg, lo, beam=W)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = {"ms": round(e0.elapsed_time(e1) / reps, 2), "launches_per_decode": (bc.launches - n0) / reps}
        outs[mode] = (bases.cpu(), nb.cpu())
        bc.check_status()
    same = all(torch.equal(outs[m][0], outs["global_kernel"][0]) and torch.equal(outs[m][1], outs["global_kernel"][1]) for m in outs)
    print(json.dumps({"B": B, "L": L, "beam": W, "bases_per_window": round(float(outs["default"][1].float().mean()), 1),
                      "modes": res, "identical": same}), flush=True)
    bc.close()
