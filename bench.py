#!/usr/bin/env python
"""bench.py -- raw-signal Msamples/s of the basecalling hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU restatement of the reference path (TF 1.15 is not installable)

A step = one pass of the hot path over one batch of synthetic windows per GPU: seq_len scaling, residual conv stack,
3-layer BiLSTM, logit head, path_prob, CTC greedy decode (DNA_default, segment_len 512, batch 4096 per GPU -- the
configuration north_star's target is quoted on).  `value` is measured with the windows resident in HBM; `e2e` goes
through the C-ABI host call (pinned host buffers, H2D + D2H inside the timed region).  Reads shard across GPUs with no
data-path collective (weak scaling)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEG_LEN = 512
FLOP_PER_FRAME = {"conv": 2 * 1049088, "lstm_in": 2 * 524800, "lstm_rec": 2 * 240000, "head": 2 * 700}
# Algorithmic HBM bytes per frame, summed over the launches of a category: every contraction reads
# its A operand image once (hi+lo fp16 = 4 B per channel) and writes its output once.
#   conv: 8 contractions read 256-channel images (the two K=512 ones read two), all write one: (10 + 8) * 1024 B, + gen 1 KB
#   lstm_in: reads 1024 + 832 + 832 B (K = 256, 208, 208), writes 3 x 3200 B of fp32 pre-activations
#   lstm_rec: reads 3 x 3200 B, writes 832 + 832 B of h images and 800 B of fp32 output;  head: 800 B in, 20 B out
BYTES_PER_FRAME = {"conv": 19 * 1024 + 4, "lstm_in": 1024 + 832 + 832 + 3 * 3200, "lstm_rec": 3 * 3200 + 832 + 832 + 800,
                   "head": 820}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synthetic_windows(B: int, L: int, seed: int):
    """Bootstrap windows cut at random offsets from the bundled normalised signals (SURVEY 8d synthetic input (i)):
    keeps the logits realistic (~95% blank) so decode costs and kbases/s are representative."""
    from chiron_b200.chiron_input import normalize_signal, read_signal
    from chiron_b200.model import NORM_UNIQUE_MAD
    rng = np.random.default_rng(seed)
    sigs = []
    for name in ("read1", "read3"):
        s = read_signal(os.path.join(ROOT, "tests", "golden", "DNA", "raw", name + ".signal"))
        sigs.append(normalize_signal(s, NORM_UNIQUE_MAD))
    x = np.empty((B, L), dtype=np.float32)
    for b in range(B):
        s = sigs[int(rng.integers(0, len(sigs)))]
        o = int(rng.integers(0, len(s) - L))
        x[b] = s[o:o + L]
    return x, np.full(B, L, dtype=np.int32)


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


def cpu_oracle_msamples(n_windows: int, L: int, seed: int, repeats: int = 1):
    """The oracle (numpy, BLAS on every host core) timed on a bounded sample of the same workload."""
    import torch
    from chiron_b200.model import load_model
    from oracle import chiron_oracle as O
    cfg, t, _ = load_model("DNA_default")
    x, lens = synthetic_windows(n_windows, L, seed)
    best = None
    bases = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        logits = O.inference(x, lens, cfg, t)
        paths = O.ctc_decode_c(logits, lens, 0)
        O.path_prob(logits)
        dt = time.perf_counter() - t0
        bases = sum(len(p) for p in paths)
        best = dt if best is None else min(best, dt)
    return n_windows * L / best / 1e6, bases / best / 1e3, best, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_windows = 96
    vals, kb = [], []
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            t_start = time.perf_counter()
        v, k, dt, threads = cpu_oracle_msamples(n_windows, SEG_LEN, 1234 + i)
        if i >= args.warmup:
            vals.append(v)
            kb.append(k)
    elapsed = time.perf_counter() - t_start
    value = n_windows * SEG_LEN * args.steps / elapsed / 1e6
    cores = os.cpu_count()
    sample = "%d windows x %d samples per step (DNA_default, greedy); numpy/BLAS oracle on all host threads" % (n_windows, SEG_LEN)
    line = {"impl": "reference", "metric": "raw-signal Msamples/s", "value": value, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "DNA_default L=512 B=4096/GPU greedy CTC (bounded CPU sample of it)", "segment_len": SEG_LEN,
                       "batch": n_windows, "decoder": "greedy"},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU restatement of the reference path (TF 1.15 cannot be installed here)"},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "kbases_per_s": float(np.mean(kb)), "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="windows per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("CHIRON_B200_PRECISION", "tc"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import ctypes
    import torch
    import torch.distributed as dist
    from chiron_b200 import _lib
    from chiron_b200.engine import Basecaller

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, L = args.batch, SEG_LEN
    bc = Basecaller("DNA_default", device=local, precision=args.precision)
    T = bc.out_len(L)
    x_h, len_h = synthetic_windows(B, L, 1234 + rank)
    dev = torch.device("cuda", local)
    x_d = torch.from_numpy(x_h).to(dev)
    len_in = torch.from_numpy(len_h).to(dev)
    len_out = torch.empty_like(len_in)
    logits = torch.empty((B, T, bc.n_class), dtype=torch.float32, device=dev)
    prob = torch.empty((B,), dtype=torch.float32, device=dev)
    bases = torch.empty((B, T), dtype=torch.int8, device=dev)
    n_bases = torch.empty((B,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def step():
        bc.seq_len_out_device(len_in, L, out=len_out, stream=stream)
        bc.forward_device(x_d, len_out, logits=logits, path_prob=prob, stream=stream)
        bc.decode_device(logits, len_out, beam=0, bases=bases, n_bases=n_bases, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = bc.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = bc.launches - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * L / ms_step / 1e3                      # Msamples/s, whole job
    total_bases = n_bases.sum().to(torch.float64)
    if world > 1:
        dist.all_reduce(total_bases)
    kbases = float(total_bases.item()) / ms_step               # bases per ms = kbases/s

    # ---- per-kernel profile of one more step (CUDA event pairs around every launch, on the launching stream) ----
    bc.enable_timing(True)
    prof_acc = {}
    n_prof = 3
    for _ in range(n_prof):
        step()
        torch.cuda.synchronize(dev)
        for k, (ms, cnt) in bc.last_forward_profile().items():
            a = prof_acc.setdefault(k, [0.0, 0])
            a[0] += ms / n_prof
            a[1] = cnt
    phase_ms = bc.last_forward_ms()
    bc.enable_timing(False)
    peaks, peak_kind = load_peaks()
    frames = B * T
    dom = max(prof_acc, key=lambda k: prof_acc[k][0])
    flops = FLOP_PER_FRAME[dom] * frames
    dom_ms, dom_cnt = prof_acc[dom]
    achieved = flops / (dom_ms * 1e-3) / 1e12
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    # tensor-pipe work actually issued: the tc path runs every contraction as 3 fp16 MMAs (hi*lo, lo*hi, hi*hi)
    mma_passes = {"tc": 3, "tc_precise": 3, "tc_fast": 1}.get(args.precision, 0)
    traffic, traffic_src = None, None
    try:                                   # DRAM bytes per launch of the dominant kernel from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tj.get("precision") == args.precision and tj.get("batch") == B and dom in tj.get("per_category", {}):
            traffic = tj["per_category"][dom]["dram_bytes_per_launch"]
            traffic_src = tj.get("source")
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "%s (bf16_tflops_sustained)" % peak_kind,
                "mma_passes": mma_passes, "tensor_pipe_tflops": achieved * max(mma_passes, 1),
                "tensor_pipe_frac": achieved * max(mma_passes, 1) / peak_tf if mma_passes else None,
                "launches_per_step": dom_cnt, "avg_launch_ms": dom_ms / max(dom_cnt, 1),
                "algorithmic_flop_per_launch": flops / max(dom_cnt, 1),
                "per_category_ms": {k: round(v[0], 3) for k, v in prof_acc.items()},
                "per_category_tflops": {k: FLOP_PER_FRAME[k] * frames / (v[0] * 1e-3) / 1e12 for k, v in prof_acc.items() if v[0] > 0},
                # every category against BOTH ceilings (tensor: FLOP actually issued = algorithmic x mma_passes vs the
                # sustained cuBLAS figure; hbm: algorithmic bytes vs the measured copy bandwidth) -- the larger is its bound
                "per_category_fracs": {k: {"tensor_pipe_frac": (FLOP_PER_FRAME[k] * frames * max(mma_passes, 1) / (v[0] * 1e-3) / 1e12 / peak_tf
                                                                if mma_passes and k != "head" else None),
                                           "hbm_frac": BYTES_PER_FRAME[k] * frames / (v[0] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                           "hbm_gbs": BYTES_PER_FRAME[k] * frames / (v[0] * 1e-3) / 1e9}
                                       for k, v in prof_acc.items() if v[0] > 0},
                "phase_ms": phase_ms}

    # ---- end to end through the C-ABI host calls: host buffers, H2D + D2H inside the timed region ---------------------
    # (a) the two-slot asynchronous pair cb_basecall_submit / cb_basecall_collect -- what chiron_eval.evaluation() drives:
    #     every step copies the batch from pinned host memory to the GPU and its result (bases, n_bases, path_prob)
    #     back; the copies of step i+1 overlap the kernels of step i.  (b) the synchronous cb_basecall_host for reference.
    lib = _lib.load()
    nbytes_x, nbytes_b = B * L * 4, B * T
    px, pl = lib.cb_host_alloc(nbytes_x), lib.cb_host_alloc(B * 4)
    pb, pn, pp = lib.cb_host_alloc(nbytes_b), lib.cb_host_alloc(B * 4), lib.cb_host_alloc(B * 4)
    if not all((px, pl, pb, pn, pp)):
        raise SystemExit("cb_host_alloc failed")
    ctypes.memmove(px, x_h.ctypes.data, nbytes_x)
    ctypes.memmove(pl, len_h.ctypes.data, B * 4)

    def e2e_sync_step():
        _lib.check(lib.cb_basecall_host(bc.h, px, pl, B, L, 0, pb, pn, pp, None), "cb_basecall_host")

    def e2e_pipelined(n_steps):
        inflight = []
        for i in range(n_steps):
            if len(inflight) == 2:
                _lib.check(lib.cb_basecall_collect(bc.h, inflight.pop(0), pb, pn, pp), "cb_basecall_collect")
            _lib.check(lib.cb_basecall_submit(bc.h, i & 1, px, pl, B, L, 0), "cb_basecall_submit")
            inflight.append(i & 1)
        while inflight:
            _lib.check(lib.cb_basecall_collect(bc.h, inflight.pop(0), pb, pn, pp), "cb_basecall_collect")

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize(dev)
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_pipelined(2)
    e2e_sync_step()
    e2e_ms = timed(lambda: e2e_pipelined(args.steps))
    sync_ms = timed(lambda: [e2e_sync_step() for _ in range(args.steps)])
    e2e_bases = int(np.ctypeslib.as_array(ctypes.cast(pn, ctypes.POINTER(ctypes.c_int32)), shape=(B,)).sum())
    e2e = {"value": world * B * L / e2e_ms / 1e3, "unit": "Msamples/s", "h2d_bytes_per_step": nbytes_x + B * 4,
           "d2h_bytes_per_step": nbytes_b + 8 * B, "ms_per_step": e2e_ms, "bases_per_step_rank0": e2e_bases,
           "api": "cb_basecall_submit/cb_basecall_collect (pinned host buffers, two batches in flight)",
           "synchronous_cb_basecall_host": {"value": world * B * L / sync_ms / 1e3, "ms_per_step": sync_ms}}
    for p in (px, pl, pb, pn, pp):
        lib.cb_host_free(p)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = 640                                             # ~10 s of CPU work on 16 threads (bounded sample)
        v, kb, dt, threads = cpu_oracle_msamples(n_cpu, L, 99)
        cpu = {"value": v, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "%d windows x %d samples of the same synthetic workload, one pass (%.1f s); numpy/BLAS oracle, %d threads"
                         % (n_cpu, L, dt, threads), "kbases_per_s": kb}
    if rank == 0:
        line = {"metric": "raw-signal Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "DNA_default L=512 B=%d/GPU greedy CTC" % B, "segment_len": L, "batch_per_gpu": B,
                           "decoder": "greedy", "precision": args.precision, "sharding": "reads/windows per rank, no collective",
                           "l2": "per-step working set (%.1f GB of activations) >> 126 MB L2; no flush needed"
                                 % (bc.lib.cb_workspace_bytes(bc.h) / 1e9),
                           "inputs": "bootstrap windows from the bundled normalised reads, seed 1234+rank"},
                "kbases_per_s": kbases, "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    bc.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
