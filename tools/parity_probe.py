"""Parity of the tensor-core mode on the bench batch (development aid; bench.py prints the same record).

    python tools/parity_probe.py [n_oracle_windows] [batch]

Runs the 4096 x 512 bench batch in `tc` and `fp32`, counts the windows whose greedy bases differ, the largest logit
difference between the two modes over the whole batch, and compares both with the float64 oracle on a sample."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bench import synthetic_windows
from chiron_b200.engine import Basecaller
from chiron_b200.model import load_model
from oracle import chiron_oracle as O

n_or = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg, t, _ = load_model("DNA_default")
x, lens = synthetic_windows(B, 512, 1234)
out = {}
for prec in ("tc", "fp32"):
    bc = Basecaller("DNA_default", device=0, precision=prec)
    out[prec] = bc.basecall_batch(x, lens, beam=0, want_logits=True)
    bc.close()
bt, nt, _, lt = out["tc"]
bf, nf, _, lf = out["fp32"]
diff = [b for b in range(B) if nt[b] != nf[b] or not np.array_equal(bt[b, :nt[b]], bf[b, :nf[b]])]
rec = {"cpp": os.environ.get("CB_TC_CPP"), "bias": os.environ.get("CB_TC_BIAS"), "windows": B, "tc_vs_fp32_mismatching_windows": len(diff),
       "tc_vs_fp32_max_dlogit": float(np.abs(lt - lf).max()), "tc_vs_fp32_rms_dlogit": float(np.sqrt(((lt - lf) ** 2).mean())),
       "argmax_flips_tc_vs_fp32": int((lt.argmax(2) != lf.argmax(2)).sum()), "frames": int(lt.shape[0] * lt.shape[1])}
if n_or > 0:
    pick = np.random.default_rng(7).choice(B, n_or, replace=False)
    ref = O.inference(x[pick], lens[pick], cfg, t, np.float64)
    paths = O.ctc_greedy(ref.astype(np.float32), lens[pick])
    am = ref.argmax(2)
    for prec, (bb, nn, _, lg) in out.items():
        rec[prec + "_vs_f64"] = {"windows": n_or, "max_dlogit": float(np.abs(lg[pick] - ref).max()),
                                 "rms_dlogit": float(np.sqrt(((lg[pick] - ref) ** 2).mean())),
                                 "argmax_flips": int((lg[pick].argmax(2) != am).sum()),
                                 "mismatching_windows": sum(bb[b, :nn[b]].tolist() != p for b, p in zip(pick, paths))}
print(json.dumps(rec))
