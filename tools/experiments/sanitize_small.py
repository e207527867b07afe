"""Tiny tc-mode forward + greedy + beam + assembly for compute-sanitizer memcheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from chiron_b200.engine import Basecaller
rng = np.random.default_rng(0)
for model, L in (("DNA_default", 256), ("RNA_default", 500)):
    bc = Basecaller(model, 0, "tc")
    x = (rng.normal(size=(130, L)) * 0.4).astype(np.float32)
    lens = np.full(130, L, np.int32); lens[5] = 17; lens[129] = 1
    b, n, p, lg = bc.basecall_batch(x, lens, beam=0, want_logits=True)
    b2, n2, p2, _ = bc.basecall_batch(x, lens, beam=20)
    t = bc.basecall_submit(0, x, lens, 0); r = bc.basecall_collect(t)
    seq, q, pos = bc.assemble(b, n, p, L - 10, L)
    print(model, "ok", int(n.sum()), int(n2.sum()), len(seq))
    bc.close()
