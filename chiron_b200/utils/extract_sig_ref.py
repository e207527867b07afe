"""fast5 -> ``<out>/raw/*.signal`` (+ ``reference/``, ``log/extract.log``): the extraction step `chiron call` runs before
basecalling (chiron/utils/extract_sig_ref.py:31-193), on top of the dependency-free HDF5 reader.

Differences from the reference, on purpose: the ``read_id`` attribute is optional (the bundled DNA fast5 files do not
carry it, which makes the reference skip them: extract_sig_ref.py:152,115-117) and ``.signal`` inputs are passed through
so a folder of already extracted signals can be basecalled with the same command."""
from __future__ import annotations

import logging
import os
import shutil
import sys
import multiprocessing
from multiprocessing import cpu_count

import numpy as np

from ..fast5 import read_fast5

logger = logging.getLogger("chiron_b200.extract")


def set_logger(log_file):
    logger.handlers = []
    handler = logging.FileHandler(log_file, mode="a")
    handler.setFormatter(logging.Formatter("%(asctime)s:%(levelname)s:%(message)s"))
    logger.addHandler(handler)
    logger.propagate = False
    logger.setLevel(logging.INFO)


def extract_file(path: str, mode: str = "dna", unit: bool = False):
    """Reads of one fast5 as [(suffix, raw_signal, read_id)] (extract_file :149-175 / extract_file_v2 :178-193)."""
    out = []
    reads = read_fast5(path)
    multi = len(reads) > 1 or (reads and not reads[0]["read_key"].startswith("Read_"))
    for r in reads:
        sig = np.asarray(r["signal"])
        if unit and r["channel"]:
            ch = r["channel"]
            sig = (sig + float(ch["offset"])) * float(ch["range"]) / float(ch["digitisation"])
        if mode == "rna":
            sig = sig[::-1]
        out.append((r["read_key"] if multi else "", sig, r["read_id"]))
    return out


def _worker(job):
    full, raw_folder, mode, unit, delimiter, idname = job
    name = os.path.basename(full)
    if full.endswith(".signal"):
        dst = os.path.join(raw_folder, name)
        if os.path.abspath(full) != os.path.abspath(dst):
            shutil.copyfile(full, dst)
        return 1
    if not full.endswith("fast5"):
        return 0
    try:
        n = 0
        for suffix, sig, read_id in extract_file(full, mode, unit):
            if len(sig) == 0:
                raise ValueError("Got empty raw signal")
            stem = read_id if (idname and read_id) else os.path.splitext(name)[0] + suffix
            with open(os.path.join(raw_folder, stem + ".signal"), "w+") as f:
                # single-read files use FLAGS.delimiter (:133-134), the multi-read branch a blank (:145-146)
                f.write((" " if suffix else delimiter).join(str(v) for v in sig.tolist()))
            n += 1
        return n
    except Exception as e:                       # extract_sig_ref.py:115-117: log and skip the file
        logger.error("Cannot extract file %s. %s" % (full, e))
        return "Cannot extract file %s. %s" % (full, e)    # the parent logs it too: a spawned worker has no log handlers


def extract(FLAGS) -> int:
    root_folder = FLAGS.input_dir
    out_folder = FLAGS.output_dir
    single = None
    if os.path.isfile(root_folder):
        single, root_folder = root_folder, os.path.dirname(os.path.abspath(root_folder))
    elif not os.path.isdir(root_folder):
        raise IOError("Input directory does not found.")
    os.makedirs(out_folder, exist_ok=True)
    FLAGS.raw_folder = os.path.abspath(os.path.join(out_folder, "raw"))
    FLAGS.ref_folder = os.path.abspath(os.path.join(out_folder, "reference"))
    FLAGS.log_folder = os.path.abspath(os.path.join(out_folder, "log"))
    for d in (FLAGS.raw_folder, FLAGS.ref_folder, FLAGS.log_folder):
        os.makedirs(d, exist_ok=True)
    set_logger(os.path.join(FLAGS.log_folder, "extract.log"))
    threads = getattr(FLAGS, "threads", 0) or cpu_count()
    jobs = []
    if single:
        files = [single]
    elif getattr(FLAGS, "recursive", True):
        files = [os.path.join(d, f) for d, _, fs in os.walk(root_folder) for f in sorted(fs)
                 if not os.path.abspath(d).startswith(os.path.abspath(out_folder) + os.sep)]
    else:
        files = [os.path.join(root_folder, f) for f in sorted(os.listdir(root_folder))]
    limit = getattr(FLAGS, "test_number", None)
    for f in files:
        if f.endswith("fast5") or f.endswith(".signal"):
            jobs.append((f, FLAGS.raw_folder, getattr(FLAGS, "mode", "dna"), getattr(FLAGS, "unit", False),
                         getattr(FLAGS, "delimiter", "\n"), getattr(FLAGS, "idname", False)))
    if limit is not None:
        jobs = jobs[:limit]
    if threads > 1 and len(jobs) > 1:
        # "spawn", not the reference's default fork (extract_sig_ref.py:83-90): by the time a long-lived caller extracts, the
        # process may hold CUDA state and the pipeline's helper threads, and forking such a process can deadlock the children
        with multiprocessing.get_context("spawn").Pool(min(threads, len(jobs))) as pool:
            counts = pool.map(_worker, jobs, chunksize=max(1, len(jobs) // (8 * threads)))
    else:
        counts = [_worker(j) for j in jobs]
    # Workers report a failure as its message: spawned workers do not inherit the parent's log handlers (the reference's
    # forked ones did), so the parent writes log/extract.log and says how many files were skipped.
    failures = [c for c in counts if isinstance(c, str)]
    if threads > 1 and len(jobs) > 1:
        for msg in failures:
            logger.error(msg)
    if failures:
        print("chiron_b200: %d of %d input files could not be extracted (see %s)"
              % (len(failures), len(jobs), os.path.join(FLAGS.log_folder, "extract.log")), file=sys.stderr)
    FLAGS.count = int(sum(c for c in counts if not isinstance(c, str)))
    FLAGS.skipped = len(failures)
    return FLAGS.count
