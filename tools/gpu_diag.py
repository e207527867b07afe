"""Per-stage error of each precision mode against the float64 oracle (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chiron_b200.engine import Basecaller
from chiron_b200.model import load_model
from oracle import chiron_oracle as O

cfg, t, _ = load_model("DNA_default")
sig = O.read_signal_text("tests/golden/DNA/raw/read1.signal")
x, lens = O.make_windows(O.normalize_signal(sig, 1), 400, 390)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
x, lens = x[40:40 + n], lens[40:40 + n]
fea64 = O.cnn_forward(x, cfg, t, np.float64)
lasth64 = O.rnn_forward(fea64, lens, cfg, t, np.float64)
lg64 = O.head_forward(lasth64, cfg, t, np.float64)
fea32 = O.cnn_forward(x, cfg, t, np.float32)
lg32 = O.inference(x, lens, cfg, t, np.float32)
print("oracle f32 vs f64: fea max %.3e rms %.3e | logits max %.3e rms %.3e | |fea|max %.1f"
      % (np.abs(fea32 - fea64).max(), np.sqrt(((fea32 - fea64) ** 2).mean()), np.abs(lg32 - lg64).max(),
         np.sqrt(((lg32 - lg64) ** 2).mean()), np.abs(fea64).max()))
am64 = lg64.argmax(2)
for prec in sys.argv[2:] or ["fp32", "tc"]:
    bc = Basecaller("DNA_default", 0, prec)
    bases, nb, prob, lg = bc.basecall_batch(x, lens, want_logits=True)
    fea = bc.debug_fetch(0, fea64.size).reshape(fea64.shape)
    lasth = bc.debug_fetch(cfg.n_layers, lasth64.size).reshape(lasth64.shape)
    flips = int((lg.argmax(2) != am64).sum())
    # relative bias of the CNN feature: least-squares slope of the error on the value (a truncating accumulator shrinks)
    slope = float(((fea - fea64) * fea64).sum() / (fea64 * fea64).sum())
    print("%-8s fea error slope %.3e (relative bias)" % (prec, slope))
    print("%-8s fea max %.3e rms %.3e | lasth max %.3e rms %.3e | logits max %.3e rms %.3e | argmax flips %d / %d"
          % (prec, np.abs(fea - fea64).max(), np.sqrt(((fea - fea64) ** 2).mean()), np.abs(lasth - lasth64).max(),
             np.sqrt(((lasth - lasth64) ** 2).mean()), np.abs(lg - lg64).max(), np.sqrt(((lg - lg64) ** 2).mean()),
             flips, am64.size))
    bc.close()
