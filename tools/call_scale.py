"""Files -> fastq throughput of a read-sharded `chiron call` on N GPUs of one box (BASELINE config 4 with files in and files out).

    python tools/call_scale.py --prepare --reads 6400 --fmt signal --dir /tmp/cs_signal          # once
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/call_scale.py --dir /tmp/cs_signal --out /tmp/cs_out [--beam 0] [--stub]

Every rank takes its share of the reads (shard.assign_reads: LPT by file size), runs chiron_eval.evaluation() on a warm
Basecaller (model, CUDA context and workspace exist before the clock starts; the ranks start together behind a barrier),
writes its own result/segments/meta files into the shared output folder, and rank 0 merges the ranks' figures exactly as
chiron_eval.run() does (meta/all.meta, meta/all.perf.json).  Rank 0 prints one JSON line: reads, samples, the wall time of
the slowest rank, whole-job Msamples/s and kbases/s, and per-rank wall/CPU seconds (CPU seconds over wall seconds = how many
host cores a rank kept busy: the host side is what bounds this path)."""
import argparse
import json
import os
import resource
import shutil
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def prepare(a):
    from chiron_b200 import chiron_input
    if a.fmt == "signal":
        src = [os.path.join(ROOT, "tests", "golden", "DNA", "raw", "read%d.signal" % i) for i in (1, 2, 3, 4, 5)]
    else:
        src = [os.path.join(ROOT, "tests", "golden", "fast5", "read1.fast5")]
    os.makedirs(a.dir, exist_ok=True)
    per = [len(chiron_input.read_signal(p)) if p.endswith(".signal") else len(chiron_input.read_signal_fast5(p)) for p in src]
    samples = 0
    for i in range(a.reads):
        j = i % len(src)
        dst = os.path.join(a.dir, "r%05d.%s" % (i, a.fmt))
        if not os.path.exists(dst):
            shutil.copy(src[j], dst)
        samples += per[j]
    with open(os.path.join(a.dir, "MANIFEST.json"), "w") as f:
        json.dump({"reads": a.reads, "samples": samples, "format": a.fmt}, f)
    print(json.dumps({"prepared": a.dir, "reads": a.reads, "samples": samples}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prepare", action="store_true")
    ap.add_argument("--reads", type=int, default=6400)
    ap.add_argument("--fmt", default="signal", choices=["signal", "fast5"])
    ap.add_argument("--dir", required=True)
    ap.add_argument("--out", default=None)
    ap.add_argument("--beam", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--stub", action="store_true", help="replace the GPU by a do-nothing Basecaller: the host pipeline alone")
    a = ap.parse_args()
    if a.prepare:
        return prepare(a)

    import torch.distributed as dist
    from chiron_b200 import chiron_eval
    from chiron_b200.shard import gather_to_rank0, rank_world
    rank, world = rank_world()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("gloo")
    with open(os.path.join(a.dir, "MANIFEST.json")) as f:
        manifest = json.load(f)
    out = a.out or (a.dir.rstrip("/") + "_out")
    if rank == 0:
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
    threads = a.threads or max(2, (os.cpu_count() or 8) // world)
    flags = types.SimpleNamespace(input=a.dir, output=out, model="DNA_default", start=None, batch_size=None, segment_len=None,
                                  jump=None, threads=threads, beam=a.beam, extension="fastq", concise=False, mode="dna",
                                  preset="dna-pre", precision=None, recursive=False)
    flags = chiron_eval.apply_preset(flags)
    if a.stub:
        from call_bench import StubCaller
        caller = StubCaller("DNA_default")
    else:
        from chiron_b200.engine import Basecaller
        caller = Basecaller("DNA_default", device=local)
        caller.basecall_batch(np.zeros((4096, flags.segment_len), np.float32), np.full(4096, flags.segment_len, np.int32), beam=a.beam)
    if world > 1:
        dist.barrier()
    r0 = resource.getrusage(resource.RUSAGE_SELF)
    t0 = time.perf_counter()
    summary = chiron_eval.evaluation(flags, caller=caller)
    wall = time.perf_counter() - t0
    r1 = resource.getrusage(resource.RUSAGE_SELF)
    time_dict = {"real": wall, "user": r1.ru_utime - r0.ru_utime, "sys": r1.ru_stime - r0.ru_stime}
    meta = os.path.join(out, "meta")
    pre = "all.rank%d" % rank if world > 1 else "all"
    chiron_eval.append_run_times(os.path.join(meta, pre + ".meta"), time_dict)
    report = chiron_eval.write_perf_report(os.path.join(meta, pre + ".perf.json"), flags, summary, time_dict, rank, world)
    parts = gather_to_rank0({"time": time_dict, "report": report})
    if rank == 0:
        if world > 1:
            merged_time = {"real": max(p["time"]["real"] for p in parts), "sys": sum(p["time"]["sys"] for p in parts),
                           "user": sum(p["time"]["user"] for p in parts)}
            chiron_eval.append_run_times(os.path.join(meta, "all.meta"), merged_time)
            merged = chiron_eval.merge_perf_reports(os.path.join(meta, "all.perf.json"), [p["report"] for p in parts])
        else:
            merged = report
        n_out = len(os.listdir(os.path.join(out, "result")))
        print(json.dumps({"what": "chiron call, files -> fastq, read-sharded%s" % (" (host pipeline only, GPU stubbed)" if a.stub else ""),
                          "n_gpus": world, "format": manifest["format"], "reads": merged["reads"], "result_files": n_out,
                          "samples": merged["samples"], "wall_s_slowest_rank": round(merged["wall_s"], 3),
                          "Msamples_per_s": round(merged["Msamples_per_s"], 2), "kbases_per_s": round(merged["kbases_per_s"], 1),
                          "precision": getattr(caller, "precision", None), "beam": a.beam, "reader_threads_per_rank": threads,
                          "host_cores": os.cpu_count(),
                          "per_rank": [{"rank": p["report"]["rank"], "reads": p["report"]["reads"], "wall_s": round(p["time"]["real"], 3),
                                        "cpu_s": round(p["time"]["user"] + p["time"]["sys"], 2)} for p in parts]}))
        assert merged["reads"] == manifest["reads"] == n_out
        shutil.rmtree(out, ignore_errors=True)
    caller.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
