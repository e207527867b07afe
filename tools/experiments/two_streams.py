"""Experiment: do two independent half-batch pipelines sharing the GPU (recurrence of one overlapping the GEMMs of the other)
beat one full-batch pipeline?  Two handles (own workspaces), two threads, two streams; aggregate Msamples/s."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from chiron_b200.engine import Basecaller

L = 512
def run(n_pipes, B, iters=12, stagger=0.0):
    bcs = [Basecaller("DNA_default", 0, "tc") for _ in range(n_pipes)]
    streams = [torch.cuda.Stream() for _ in range(n_pipes)]
    xs = [torch.randn(B, L, device="cuda") * 0.43 - 0.16 for _ in range(n_pipes)]
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    def work(i, n):
        with torch.cuda.stream(streams[i]):
            for _ in range(n):
                lg, pr = bcs[i].forward_device(xs[i], lens, stream=streams[i].cuda_stream)
                bcs[i].decode_device(lg, lens, stream=streams[i].cuda_stream)
    for i in range(n_pipes): work(i, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = []
    for i in range(n_pipes):
        t = threading.Thread(target=work, args=(i, iters)); t.start(); th.append(t)
        if stagger: time.sleep(stagger)
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for b in bcs: b.close()
    return n_pipes * B * L * iters / dt / 1e6

print("1 pipe  x 4096: %.1f Msamples/s" % run(1, 4096))
print("2 pipes x 2048: %.1f Msamples/s" % run(2, 2048))
print("2 pipes x 2048 staggered 8 ms: %.1f Msamples/s" % run(2, 2048, stagger=0.008))
print("2 pipes x 4096 staggered 15 ms: %.1f Msamples/s" % run(2, 4096, stagger=0.015))
print("4 pipes x 1024 staggered 4 ms: %.1f Msamples/s" % run(4, 1024, stagger=0.004))
