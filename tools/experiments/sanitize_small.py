"""Tiny tc-mode forward + greedy + beam + assembly for compute-sanitizer memcheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from chiron_b200.engine import Basecaller
rng = np.random.default_rng(0)
for model, L in (("DNA_default", 256), ("RNA_default", 500)):
    bc = Basecaller(model, 0, "tc")
    x = (rng.normal(size=(300, L)) * 0.4).astype(np.float32)
    lens = np.full(300, L, np.int32); lens[5] = 17; lens[129] = 1
    b, n, p, lg = bc.basecall_batch(x, lens, beam=0, want_logits=True)
    b2, n2, p2, _ = bc.basecall_batch(x, lens, beam=20)
    t = bc.basecall_submit(0, x, lens, 0); r = bc.basecall_collect(t)
    seq, q, pos = bc.assemble(b, n, p, L - 10, L)
    out = bc.predict(x[:40], lens[:40], beam_width=20)            # cb_decode_beam_scored
    os.environ["CB_BEAM_POOL"] = str(2 * 20 + 2)                  # force the retry / fallback passes of the beam search
    b3, n3, _, _ = bc.basecall_batch(x, lens, beam=20)
    del os.environ["CB_BEAM_POOL"]
    assert np.array_equal(n2, n3) and np.array_equal(b2, b3)
    print(model, "ok", int(n.sum()), int(n2.sum()), len(seq), out["log_prob"].shape)
    bc.close()
