"""Pipeline driver: the drop-in for chiron/chiron_eval.py (``run(args)`` / ``evaluation()``).

Same flags, same output tree (``result/ segments/ meta/``), same file formats (write_output, chiron_eval.py:176-242);
the TF session, the logits FIFO queue, the six decode threads and the pure-Python assembly are replaced by calls into
libchiron_b200.so (forward + CTC decode + assembly on the GPU).  Differences that are deliberate, documented in
DESIGN.md:
  * windows of a read are collated in TRUE window order (the reference keys chunks by batch position,
    chiron_eval.py:413-428,436, which rotates/duplicates segments of multi-batch reads -- SURVEY.md finding 6);
  * the per-model signal normalisation the shipped weights need is applied (SURVEY.md finding 4);
  * reads shard across ranks when launched under torchrun (RANK/WORLD_SIZE); no collective on the decode path.

Host pipeline (replaces the 1 feeder + 6 decode threads + FIFO queue of chiron_eval.py:262-268,495-521):
reader threads parse / normalise / window the next reads while the GPU works; batches go through the two-slot
asynchronous C-ABI call (pinned staging, copy-in / compute / copy-out streams), so the copies and the host-side
collation of batch i overlap the kernels of batch i+1; a writer thread formats and writes the output files.
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import queue
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional

import numpy as np

from .chiron_input import read_data_for_eval
from ._lib import BN_BATCH
from .engine import Basecaller, format_segments, get_assembler_kernal, index2base
from .shard import assign_reads, gather_to_rank0, rank_world
from .utils.unix_time import unix_time

FLAGS = None


# File formats of the output tree (chiron_eval.py:176-242).  Only these byte layouts are shared with the reference; the
# records of a segments file are FASTA-style even under a .fastq suffix (the reference never passes seg_q_score, :213-215,461)
# and a fasta result has no trailing newline (:220).
_FMT = {
    "segment": ">{name}{idx}\n{seq}\n",
    "segment_q": "@{name}{idx}\n{seq}\n+\n{qual}\n",
    "result_fastq": "@{name}\n{seq}\n+\n{qual}\n",
    "result_fasta": ">{name}\n{seq}",
    "meta": ("# Reading Basecalling assembly output total rate(bp/s)\n"
             "{reading:5.3f} {basecall:5.3f} {assembly:5.3f} {output:5.3f} {total:5.3f} {rate:5.3f}\n"
             "# read_len batch_size segment_len jump start_pos\n"
             "{read_len:d} {batch_size:d} {segment_len:d} {jump:d} {start:d}\n"
             "# input_name model_name\n"
             "{input} {model}\n"),
}


def _segment_records(file_pre: str, segments, with_q: bool, seg_q_score) -> str:
    parts = []
    for idx, seq in enumerate(segments):
        parts.append(_FMT["segment"].format(name=file_pre, idx=idx, seq=seq))
        if with_q:
            parts.append(_FMT["segment_q"].format(name=file_pre, idx=idx, seq=seq, qual=seg_q_score[idx]))
    return "".join(parts)


def write_output(segments, consensus: str, time_list, file_pre: str, global_setting, concise: bool = False,
                 suffix: str = "fasta", seg_q_score=None, q_score: Optional[str] = None):
    """The three files of one read, byte-compatible with chiron_eval.py:176-242 (pinned by the seven reference-run cases of
    tests/golden/host_ref).  ``segments`` is the reference's list of base strings, or the already formatted records as
    ``bytes`` (engine.format_segments: what evaluation() passes).  ``time_list`` holds the cumulative stamps
    [start, reading, basecalling, assembly] the reference keeps; the meta line reports their successive differences."""
    out_dir = global_setting.output
    if global_setting.mode == "rna":
        consensus = consensus.replace("T", "U").replace("t", "u")
    fastq = suffix == "fastq"
    result = (_FMT["result_fastq"].format(name=file_pre, seq=consensus, qual=q_score) if fastq and q_score is not None
              else _FMT["result_fasta"].format(name=file_pre, seq=consensus))
    with open(os.path.join(out_dir, "result", file_pre + "." + suffix), "w") as f:
        f.write(result)
    if concise:
        return
    if not isinstance(segments, (bytes, bytearray)):
        segments = _segment_records(file_pre, segments, fastq and seg_q_score is not None, seg_q_score).encode("utf-8")
    with open(os.path.join(out_dir, "segments", file_pre + "." + suffix), "wb") as f:
        f.write(segments)
    start, reading, basecall_cum, assembly_cum = time_list
    output = (time.time() - start) - assembly_cum           # everything after the assembly stamp: formatting + writing
    total = time.time() - start
    meta = _FMT["meta"].format(reading=reading, basecall=basecall_cum - reading, assembly=assembly_cum - basecall_cum,
                               output=output, total=total, rate=len(consensus) / total, read_len=len(consensus),
                               batch_size=global_setting.batch_size, segment_len=global_setting.segment_len,
                               jump=global_setting.jump, start=global_setting.start, input=global_setting.input,
                               model=global_setting.model)
    with open(os.path.join(out_dir, "meta", file_pre + ".meta"), "w") as f:
        f.write(meta)


def list_input_files(flags) -> (List[str], str):
    """chiron_eval.py:277-290.  With -r the reference builds the names of files in sub-folders without a path separator
    (``dirpath[dir_len:] + filename``, :283), so its recursive mode only works on flat folders; here a nested read keeps
    its relative path as the input name and gets a flattened output prefix (see ``output_prefix``)."""
    if os.path.isdir(flags.input):
        if getattr(flags, "recursive", False):
            file_list = []
            for (dirpath, dirnames, filenames) in os.walk(flags.input):
                dirnames.sort()
                for filename in sorted(filenames):
                    file_list.append(os.path.relpath(os.path.join(dirpath, filename), flags.input))
        else:
            file_list = sorted(os.listdir(flags.input))
        file_dir = flags.input
    else:
        file_list = [os.path.basename(flags.input)]
        file_dir = os.path.abspath(os.path.join(flags.input, os.path.pardir))
    return [f for f in file_list if f.endswith(".signal") or f.endswith(".fast5")], file_dir


def output_prefix(name: str) -> str:
    """result/ segments/ meta/ file prefix of an input name: the name without its extension; the path separators of a
    read found in a sub-folder (-r) become '__', so every read lands in the flat output folders the reference writes."""
    return os.path.splitext(name)[0].replace(os.sep, "__")


class _ReadState:
    __slots__ = ("name", "n", "samples", "bases", "n_bases", "prob", "filled", "start_time", "reading_time")

    def __init__(self, name, n, T, start_time, reading_time, samples=0):
        self.name = name
        self.n = n
        self.samples = samples
        self.bases = np.zeros((n, T), dtype=np.int8)
        self.n_bases = np.zeros(n, dtype=np.int32)
        self.prob = np.zeros(n, dtype=np.float32)
        self.filled = 0
        self.start_time = start_time
        self.reading_time = reading_time


def evaluation(flags=None, caller: Optional[Basecaller] = None) -> Dict[str, dict]:
    """chiron_eval.py:378-463.  Windows are packed ACROSS reads into batches of ``batch_size`` exactly like
    ``_worker_fn`` (:304-368); the final partial batch is run at its true size (with population BatchNorm a window's
    result does not depend on the other rows, so the reference's wrap-padding rows are dead work)."""
    flags = flags or FLAGS
    rank, world = rank_world()
    own = caller is None
    if own:
        device = int(os.environ.get("LOCAL_RANK", getattr(flags, "device", 0) or 0))
        caller = Basecaller(flags.model, device=device, precision=getattr(flags, "precision", None) or "auto")
    cfg = caller.cfg
    try:
        flags.precision_used = getattr(caller, "precision", None)      # what "auto" resolved to (perf report)
    except AttributeError:
        pass
    file_list, file_dir = list_input_files(flags)
    if world > 1:
        sizes = [os.path.getsize(os.path.join(file_dir, f)) for f in file_list]
        file_list = [file_list[i] for i in assign_reads(sizes, world)[rank]]
    print("Found %d files." % len(file_list))
    for sub in ("", "segments", "result", "meta"):
        os.makedirs(os.path.join(flags.output, sub), exist_ok=True)
    L, jump, B = flags.segment_len, flags.jump, flags.batch_size
    # GPU batch: a window's result does not depend on the batch it sits in (population BatchNorm; pinned bit-exactly by
    # tests/test_gpu_forward.py::test_full_size_batch_properties), and the recurrence is a latency chain that costs the
    # same for 400 windows as for 4096, so windows are packed into batches of at least `gpu_batch` whatever -b says
    # (the presets' 300-400 windows would leave 90 % of the SMs idle).  The meta files still record flags.batch_size.
    gpu_batch = int(os.environ.get("CHIRON_B200_GPU_BATCH", min(4096, max(1, (4 << 20) // max(L, 1)))))
    # Batch-statistics BatchNorm (models trained at HEAD, chiron/cnn.py:166-188): a window's result DOES depend on the
    # batch it sits in, so the batches must be the reference's own -- exactly -b windows each, the final one wrap-padded
    # (chiron_eval.py:322-329,352-358); the padded rows take part in the batch moments and are then discarded.
    batch_bn = caller.bn_mode == BN_BATCH
    if not batch_bn:
        B = max(B, gpu_batch)
    T = caller.out_len(L)
    beam = flags.beam
    with_qs = flags.extension == "fastq"
    summary: Dict[str, dict] = {}

    pend_x: List[np.ndarray] = []
    pend_len: List[np.ndarray] = []
    pend_owner: List[tuple] = []      # (state, first window index, count)
    pend_n = 0
    open_reads: List[_ReadState] = []

    def assemble_read(st: _ReadState):
        basecall_time = time.time() - st.start_time
        kernal = get_assembler_kernal(jump, L)
        if os.environ.get("CHIRON_B200_SKIP_ASSEMBLY") == "1":         # development: what the per-read assembly costs the run
            seq, qual, pos = "", ("" if with_qs else None), np.zeros(st.n, np.int32)
        else:
            seq, qual, pos = caller.assemble(st.bases, st.n_bases, st.prob if with_qs else None, jump, L, kernel=kernal,
                                             with_qs=with_qs)
        assembly_time = time.time() - st.start_time
        file_pre = output_prefix(st.name)
        # the segment strings are built by the writer thread: this thread's job is to keep the GPU fed
        write_q.put((st.bases, st.n_bases, seq, [st.start_time, st.reading_time, basecall_time, assembly_time], file_pre, qual))
        summary[st.name] = {"windows": st.n, "samples": st.samples, "bases": len(seq), "pos": pos, "kept": st.n_bases > 0}

    # ---- finisher thread: per-read assembly off the GPU-feeding thread ---------------------------------------------------
    # cb_assemble_host synchronises its own (default) stream; while the forward kernels of two batches occupy the SMs the
    # small assembly kernels of a read can wait milliseconds for a slot, and a thread that waited for them could not submit
    # the next batch.  The assembly staging of a handle is separate from the submit/collect slots, so one finisher thread
    # may run cb_assemble_host while the main thread runs cb_basecall_submit / cb_basecall_collect (include/chiron_b200.h).
    # CHIRON_B200_FINISHER=0 assembles on the main thread instead.
    use_finisher = os.environ.get("CHIRON_B200_FINISHER", "1") != "0"
    finish_q: "queue.Queue" = queue.Queue()
    finish_err: List[BaseException] = []

    def finisher():
        while True:
            st = finish_q.get()
            if st is None:
                return
            try:
                if not finish_err:
                    assemble_read(st)
            except BaseException as e:            # surfaced after the join below
                finish_err.append(e)

    finisher_thread = threading.Thread(target=finisher, name="chiron-finisher", daemon=True)
    if use_finisher:
        # the assembly kernels of finished reads run next to the forward kernels of the following batches: keep a few SMs
        # out of the persistent contraction grids so that they find a slot (cb_reserve_sms; +7 % on files -> fastq)
        if hasattr(caller, "reserve_sms"):
            caller.reserve_sms(int(os.environ.get("CHIRON_B200_RESERVE_SMS", "4")))
        finisher_thread.start()

    def finish(st: _ReadState):
        if use_finisher:
            finish_q.put(st)
        else:
            assemble_read(st)

    # ---- writer threads: formatting and file I/O off the GPU-feeding thread (one file set per read: order-free) ----------
    write_q: "queue.Queue" = queue.Queue()
    write_err: List[BaseException] = []

    def writer():
        while True:
            item = write_q.get()
            if item is None:
                return
            try:
                bases, n_bases, seq, times, file_pre, qual = item
                records = format_segments(file_pre, bases, n_bases)
                write_output(records, seq, times, file_pre, concise=flags.concise, suffix=flags.extension, q_score=qual,
                             global_setting=flags)
            except BaseException as e:            # surfaced after the joins below
                write_err.append(e)

    n_writers = max(1, min(4, (os.cpu_count() or 2) // 4))
    writer_threads = [threading.Thread(target=writer, name="chiron-writer-%d" % i, daemon=True) for i in range(n_writers)]
    for t in writer_threads:
        t.start()

    # ---- two batches in flight on the GPU (slots 0/1 of cb_basecall_submit) ----------------------------------------------
    inflight = collections.deque()               # (ticket, owners) in submission order

    def collect_oldest():
        ticket, owners = inflight.popleft()
        bases, n_bases, prob = caller.basecall_collect(ticket)
        off = 0
        for st, first, cnt in owners:
            st.bases[first:first + cnt] = bases[off:off + cnt]
            st.n_bases[first:first + cnt] = n_bases[off:off + cnt]
            st.prob[first:first + cnt] = prob[off:off + cnt]
            st.filled += cnt
            off += cnt
        while open_reads and open_reads[0].filled == open_reads[0].n:
            finish(open_reads.pop(0))

    next_slot = 0

    def flush():
        nonlocal pend_x, pend_len, pend_owner, pend_n, next_slot
        if pend_n == 0:
            return
        x = np.concatenate(pend_x, axis=0)
        ln = np.concatenate(pend_len, axis=0)
        if batch_bn and pend_n < B:               # only the final batch can be short (chiron_eval.py:352-358)
            x = np.pad(x, ((0, B - pend_n), (0, 0)), mode="wrap")
            ln = np.pad(ln, (0, B - pend_n), mode="wrap")
        if len(inflight) == 2:                    # the slot about to be reused is the oldest batch
            collect_oldest()
        inflight.append((caller.basecall_submit(next_slot, x, ln, beam=beam), pend_owner))
        next_slot ^= 1
        pend_x, pend_len, pend_owner, pend_n = [], [], [], 0

    # ---- reader threads: parse / normalise / window ahead of the GPU (results consumed in file order) -------------------
    n_readers = getattr(flags, "threads", 0) or min(8, os.cpu_count() or 1)

    def load(name):
        t0 = time.time()
        data = read_data_for_eval(os.path.join(file_dir, name), flags.start, seg_length=L, step=jump,
                                  reverse_fast5=getattr(flags, "reverse_fast5", False), sig_norm=cfg.sig_norm)
        return data, t0, time.time() - t0

    pool = ThreadPoolExecutor(max_workers=n_readers, thread_name_prefix="chiron-reader")
    lookahead = 2 * n_readers
    failed = True
    try:
        futures = collections.deque(pool.submit(load, n) for n in file_list[:lookahead])
        for idx, name in enumerate(file_list):
            eval_data, start_time, reading_time = futures.popleft().result()
            if idx + lookahead < len(file_list):
                futures.append(pool.submit(load, file_list[idx + lookahead]))
            st = _ReadState(name, eval_data.reads_n, T, start_time, reading_time, getattr(eval_data, "samples", 0))
            open_reads.append(st)
            i = 0
            if eval_data.reads_n == 0:
                continue
            while eval_data.epochs_completed == 0:
                cur_x, cur_len, _ = eval_data.next_batch(B - pend_n, shuffle=False)
                cnt = len(cur_x)
                pend_x.append(cur_x)
                pend_len.append(cur_len)
                pend_owner.append((st, i, cnt))
                pend_n += cnt
                i += cnt
                if pend_n >= B:
                    flush()
        flush()
        while inflight:
            collect_oldest()
        while open_reads:                                            # reads with zero windows
            finish(open_reads.pop(0))
        failed = False
    finally:
        # Wind the helper threads down on every path: an exception (unreadable input, a CUDA error) must surface with no
        # reader / finisher / writer thread and no GPU handle left behind when evaluation() is used as a library call.
        pool.shutdown(wait=True, cancel_futures=True)
        if failed:
            while inflight:                                          # slots must be collected before the handle goes
                try:
                    caller.basecall_collect(inflight.popleft()[0])
                except Exception:
                    pass
        if use_finisher:
            finish_q.put(None)
            finisher_thread.join()
        for _ in writer_threads:
            write_q.put(None)
        for t in writer_threads:
            t.join()
        if own:
            caller.close()
    if finish_err:
        raise finish_err[0]
    if write_err:
        raise write_err[0]
    return summary


def run(args):
    """chiron_eval.py:525-544."""
    global FLAGS
    FLAGS = args
    print("The result will be written to %s" % (FLAGS.output))
    os.makedirs(FLAGS.output, exist_ok=True)
    result = {}
    time_dict = unix_time(lambda: result.update(evaluation()))
    print("Real time:%5.3f Systime:%5.3f Usertime:%5.3f" % (time_dict["real"], time_dict["sys"], time_dict["user"]))
    meta_folder = os.path.join(FLAGS.output, "meta")
    if os.path.isdir(FLAGS.input):
        file_pre = "all"
    else:
        file_pre = os.path.splitext(os.path.basename(FLAGS.input))[0]
    rank, world = rank_world()
    rank_pre = file_pre + ".rank%d" % rank if world > 1 else file_pre
    append_run_times(os.path.join(meta_folder, rank_pre + ".meta"), time_dict)
    report = write_perf_report(os.path.join(meta_folder, rank_pre + ".perf.json"), FLAGS, result, time_dict, rank, world)
    if world > 1:
        # read-sharded run: rank 0 merges the ranks' figures into the files a single-process run writes:
        # wall time = the slowest rank, CPU times and counts = sums
        parts = gather_to_rank0({"time": time_dict, "report": report})
        if rank == 0:
            merged_time = {"real": max(p["time"]["real"] for p in parts), "sys": sum(p["time"]["sys"] for p in parts),
                           "user": sum(p["time"]["user"] for p in parts)}
            append_run_times(os.path.join(meta_folder, file_pre + ".meta"), merged_time)
            merge_perf_reports(os.path.join(meta_folder, file_pre + ".perf.json"), [p["report"] for p in parts])


def append_run_times(path_meta: str, time_dict: dict):
    """The run's line pair of meta/all.meta (chiron_eval.py:541-544; appended, like the reference)."""
    with open(path_meta, "a+") as out_meta:
        out_meta.write("# Wall_time Sys_time User_time Cpu_time\n")
        out_meta.write("%5.3f %5.3f %5.3f %5.3f\n" % (time_dict["real"], time_dict["sys"], time_dict["user"],
                                                      time_dict["sys"] + time_dict["user"]))


def merge_perf_reports(path: str, reports: List[dict]) -> dict:
    """all.perf.json of a read-sharded run: counts and CPU seconds summed over the ranks, wall time of the slowest."""
    merged = dict(reports[0])
    for k in ("reads", "windows", "samples", "bases", "cpu_s"):
        merged[k] = sum(r[k] for r in reports)
    merged["wall_s"] = max(r["wall_s"] for r in reports)
    wall = max(merged["wall_s"], 1e-9)
    merged["Msamples_per_s"] = merged["samples"] / wall / 1e6
    merged["kbases_per_s"] = merged["bases"] / wall / 1e3
    merged["rank"] = "merged"
    merged["per_rank"] = [{k: r[k] for k in ("rank", "reads", "samples", "bases", "wall_s", "Msamples_per_s")} for r in reports]
    with open(path, "w") as f:
        json.dump(merged, f, indent=1)
        f.write("\n")
    return merged


def write_perf_report(path: str, flags, summary: Dict[str, dict], time_dict: dict, rank: int = 0, world: int = 1) -> dict:
    """Machine-readable twin of meta/all.meta ("add JSON perf report"): what this rank processed and
    the raw-signal Msamples/s / kbases/s it delivered end to end, files in to files out.  The reference has no such file;
    it sits next to all.meta and changes nothing else in the output tree."""
    samples = sum(v.get("samples", 0) for v in summary.values())
    bases = sum(v["bases"] for v in summary.values())
    wall = max(time_dict["real"], 1e-9)
    report = {"reads": len(summary), "windows": sum(v["windows"] for v in summary.values()), "samples": samples, "bases": bases,
              "wall_s": time_dict["real"], "cpu_s": time_dict["sys"] + time_dict["user"],
              "Msamples_per_s": samples / wall / 1e6, "kbases_per_s": bases / wall / 1e3,
              "segment_len": flags.segment_len, "jump": flags.jump, "batch_size": flags.batch_size, "beam": flags.beam,
              "precision": getattr(flags, "precision_used", None) or getattr(flags, "precision", None) or "auto",
              "model": flags.model, "mode": flags.mode,
              "rank": rank, "world_size": world}
    with open(path, "w") as f:
        json.dump(report, f, indent=1)
        f.write("\n")
    return report


def add_call_arguments(parser: argparse.ArgumentParser, model_default: Optional[str] = None):
    """The flag surface shared by `chiron call` (entry.py:69-92) and the stand-alone driver (chiron_eval.py:546-575)."""
    parser.add_argument("-i", "--input", required=True, help="File path or Folder path to the fast5 / signal files.")
    parser.add_argument("-o", "--output", required=True, help="Output Folder name")
    parser.add_argument("-m", "--model", required=model_default is None, default=model_default, help="model folder path")
    parser.add_argument("-s", "--start", type=int, default=None, help="Start index of the signal file.")
    parser.add_argument("-b", "--batch_size", type=int, default=None, help="Batch size for run.")
    parser.add_argument("-l", "--segment_len", type=int, default=None, help="Segment length to be divided into.")
    parser.add_argument("-j", "--jump", type=int, default=None, help="Step size for segment")
    parser.add_argument("-t", "--threads", type=int, default=None, help="Threads number (host-side parsing only)")
    parser.add_argument("--beam", type=int, default=None,
                        help="Beam width of the beam search decoder; 0 = greedy decoder.")
    parser.add_argument("-e", "--extension", default="fastq", help="Output file extension.")
    parser.add_argument("--concise", action="store_true", help="Do not write the meta and segments files.")
    parser.add_argument("--mode", default="dna", help="Output mode, dna or rna.")
    parser.add_argument("-p", "--preset", default=None, help="Preset evaluation parameters: dna-pre or rna-pre")
    parser.add_argument("--precision", default="auto", choices=["auto", "tc", "fp32"],
                        help="[chiron_b200] arithmetic of the dense contractions (see include/chiron_b200.h): tc = the "
                             "tcgen05 tensor-core kernels, fp32 = the FFMA kernels, auto = tc whenever the model's "
                             "topology is covered by them")


def apply_preset(args):
    """entry.py:19-31,53-60 / chiron_eval.py:577-600."""
    if args.preset is None:
        p = {"start": 0, "batch_size": 400, "segment_len": 500, "jump": 490, "threads": 0, "beam": 30}
    elif args.preset == "dna-pre":
        p = {"start": 0, "batch_size": 400, "segment_len": 400, "jump": 390, "threads": 0, "beam": 30}
        if args.mode == "rna":
            raise ValueError("Try to use the DNA preset parameter setting in RNA mode.")
    elif args.preset == "rna-pre":
        p = {"start": 0, "batch_size": 300, "segment_len": 2000, "jump": 1900, "threads": 0, "beam": 30}
        if args.mode == "dna":
            raise ValueError("Attempt to use the RNA preset parameter setting in DNA mode, enable rna mode by --mode.")
    else:
        raise ValueError("Unknown presetting %s undifiend" % (args.preset))
    for k, v in p.items():
        if getattr(args, k, None) is None:
            setattr(args, k, v)
    args.reverse_fast5 = args.mode == "rna"
    return args


def main(argv=None):
    parser = argparse.ArgumentParser(prog="chiron", description="A deep neural network basecaller (B200-native path).")
    add_call_arguments(parser)
    parser.add_argument("-r", "--recursive", action="store_true", help="If read the files recursively.")
    args = parser.parse_args(sys.argv[1:] if argv is None else argv)
    run(apply_preset(args))


if __name__ == "__main__":
    main()
