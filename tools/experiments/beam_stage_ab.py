"""Same-process A/B of beam_warp_kernel with the window's logits staged in shared memory (CB_BEAM_STAGE_LOGITS=1, the earlier
variant) against read from global memory with a one-frame prefetch (=0: shared memory holds the workspace only, so 5 instead
of 2 CTAs are resident per SM at T=512, W=30).  Prints one JSON line per case; outputs must be identical."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from chiron_b200.engine import Basecaller

for model, B, L, W in (("DNA_default", 4096, 512, 30), ("DNA_default", 400, 400, 30), ("RNA_default", 512, 500, 50)):
    bc = Basecaller(model, 0, "tc")
    x = torch.randn(B, L, device="cuda") * 0.43 - 0.16
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    lo = bc.seq_len_out_device(lens, L)
    lg, _ = bc.forward_device(x, lo)
    torch.cuda.synchronize()
    res = {}
    for staged in ("1", "0", "1", "0"):
        os.environ["CB_BEAM_STAGE_LOGITS"] = staged
        bases, nb = bc.decode_device(lg, lo, beam=W)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            bases, nb = bc.decode_device(lg, lo, beam=W)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        res.setdefault(staged, []).append(round(ms, 3))
        res["sum" + staged] = (int(bases.to(torch.int64).sum()), int(nb.sum()))
    print(json.dumps({"model": model, "B": B, "L": L, "beam": W, "staged_ms": res["1"], "global_prefetch_ms": res["0"],
                      "identical": res["sum0"] == res["sum1"]}), flush=True)
    bc.close()
