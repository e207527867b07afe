"""Quick on-GPU timing of cb_forward phases (development aid, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chiron_b200.engine import Basecaller

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
shapes = [(1024, 512), (4096, 512)] if len(sys.argv) < 4 else [(int(sys.argv[2]), int(sys.argv[3]))]
bc = Basecaller("DNA_default", 0, prec)
bc.enable_timing(True)
for B, L in shapes:
    x = torch.randn(B, L, device="cuda") * 0.43 - 0.16
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    for it in range(3):
        logits, prob = bc.forward_device(x, lens)
        bases, nb = bc.decode_device(logits, lens)
        torch.cuda.synchronize()
        ms = bc.last_forward_ms()
        print(prec, B, L, "ms conv/lstm/head/total", [round(v, 2) for v in ms],
              "Msamples/s %.2f" % (B * L / ms[3] / 1e3), "mean bases/window %.1f" % nb.float().mean().item(), flush=True)
if os.environ.get("CB_PROF_DUMP"):
    bc.last_forward_profile()
