"""Time the fp32 forward in both BatchNorm modes (batch-statistics BN = HEAD's simple_global_bn).

    python tools/bn_bench.py [B] [L] [iters]

Prints one JSON line: ms per forward and launches per forward in population and batch mode.  Under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:bn_` it gives the achieved HBM
bandwidth of the bn_* kernels (tools/ncu_bn_summary.py)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from chiron_b200.engine import Basecaller


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.normal(-0.16, 0.43, size=(B, L)).astype(np.float32)).cuda()
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    out = {"B": B, "L": L, "iters": iters}
    for mode in ("population", "batch"):
        bc = Basecaller("DNA_default", device=0, precision="fp32", bn_mode=mode)
        bc.enable_timing(True)
        bc.forward_device(x, lens)                      # warm-up (workspace allocation)
        torch.cuda.synchronize()
        n0 = bc.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            bc.forward_device(x, lens)
        e1.record()
        torch.cuda.synchronize()
        out[mode] = {"ms_per_forward": e0.elapsed_time(e1) / iters, "launches_per_forward": (bc.launches - n0) // iters,
                     "phase_ms": bc.last_forward_ms()}
        bc.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
