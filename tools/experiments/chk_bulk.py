import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import chiron_oracle as O
from chiron_b200.model import load_model
from chiron_b200.engine import Basecaller
cfg, t, _ = load_model("DNA_default")
sig = O.read_signal_text("/root/repo/tests/golden/DNA/raw/read1.signal")
x, lens = O.make_windows(O.normalize_signal(sig, cfg.sig_norm), 200, 150)
x, lens = x[:150].copy(), lens[:150].copy()
ref = O.inference(x, lens, cfg, t, np.float64)
bc = Basecaller("DNA_default", 0, "tc")
bases, nb, prob, lg = bc.basecall_batch(x, lens, beam=0, want_logits=True)
err = np.abs(lg - ref)
print(os.environ.get("CHIRON_B200_LIB", "working tree"), "max err", err.max(), "rows>=128 max", err[128:].max(), "rows<128 max", err[:128].max(), "argmax flips", int((lg.argmax(2) != ref.argmax(2)).sum()))
