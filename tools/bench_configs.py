"""BASELINE.json configs 2, 3 and 5 on one B200 (bench.py carries config 4's per-GPU workload):
  config 2: DNA_default, x[1024,512], greedy;   config 3: RNA_default, x[512,500], beam 50;
  config 5: sweep segment_len {300,512,1024,2048} x batch {256..8192} (B*L <= 4 Mi frames), per-category ms.
Prints one JSON document (device-resident inputs, CUDA events on the launching stream, 3 warm-up + 5 timed passes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_windows
from chiron_b200.engine import Basecaller

def run(bc, B, L, beam, iters=5, warm=3, model="DNA_default"):
    xs, lens_h = synthetic_windows(B, L, 1234, model)       # bootstrap windows of the bundled reads
    x, lens = torch.from_numpy(xs).cuda(), torch.from_numpy(lens_h).cuda()
    lo = bc.seq_len_out_device(lens, L)
    def step():
        logits, prob = bc.forward_device(x, lo)
        return bc.decode_device(logits, lo, beam=beam)
    for _ in range(warm): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): bases, nb = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bc.enable_timing(True); step(); torch.cuda.synchronize()
    prof = {k: round(v[0], 3) for k, v in bc.last_forward_profile().items()}
    bc.enable_timing(False)
    return {"batch": B, "segment_len": L, "beam": beam, "ms_per_pass": round(ms, 3), "Msamples_per_s": round(B * L / ms / 1e3, 2),
            "windows_per_s": round(B / ms * 1e3, 1), "per_category_ms": prof}

prec = sys.argv[1] if len(sys.argv) > 1 else "tc"
out = {"precision": prec, "inputs": "bootstrap windows cut from the bundled normalised reads (seed 1234), full windows, resident in HBM"}
dna = Basecaller("DNA_default", 0, prec)
out["config2_dna_1024x512_greedy"] = run(dna, 1024, 512, 0)
sweep = []
for L in (300, 512, 1024, 2048):
    for B in (256, 1024, 4096, 8192):
        if B * L > 4 * 1024 * 1024: continue
        sweep.append(run(dna, B, L, 0, iters=3, warm=2))
        print("sweep", sweep[-1], file=sys.stderr, flush=True)
out["config5_sweep_dna_greedy"] = sweep
dna.close()
rna = Basecaller("RNA_default", 0, prec)
out["config3_rna_512x500_beam50"] = run(rna, 512, 500, 50, model="RNA_default")
out["config3_rna_512x500_greedy"] = run(rna, 512, 500, 0, model="RNA_default")
rna.close()
print(json.dumps(out))
