"""CPU tests of the host-side I/O mirror: HDF5/fast5 reader, extraction, normalisation, windowing, sharding."""
import os
import time
import types

import numpy as np
import pytest

from conftest import GOLDEN
from chiron_b200 import chiron_input, fast5, shard
from chiron_b200.chiron_eval import apply_preset, list_input_files
from chiron_b200.engine import get_assembler_kernal
from chiron_b200.model import NORM_FULL_MAD, NORM_UNIQUE_MAD
from chiron_b200.utils.extract_sig_ref import extract
from oracle import chiron_oracle as O

DNA_FAST5 = os.path.join(GOLDEN, "fast5", "read1.fast5")
RNA_FAST5 = os.path.join(GOLDEN, "fast5", "rna_read_100_ch_328.fast5")


def test_dna_fast5_equals_golden_signal():
    """Old-style groups + contiguous int16 dataset; the payload equals output/raw/read1.signal sample for sample."""
    sig = fast5.read_raw_signal(DNA_FAST5)
    ref = chiron_input.read_signal(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    assert sig.dtype == np.int16 and len(sig) == 62461
    assert np.array_equal(sig.astype(np.float32), ref)
    rec = fast5.read_fast5(DNA_FAST5)[0]
    assert rec["read_id"] is None                       # the bundled DNA reads carry no read_id attribute
    assert rec["attrs"]["duration"] == 62461 and rec["channel"]["sampling_rate"] == 4000.0


def test_rna_fast5_chunked_deflate_and_link_messages():
    """New-style compact groups + chunked, deflate-compressed Signal."""
    rec = fast5.read_fast5(RNA_FAST5)[0]
    assert rec["read_key"] == "Read_100"
    assert rec["read_id"] == "ed5cc40d-a190-4ed2-8c2c-02d89e092c59"
    assert len(rec["signal"]) == rec["attrs"]["duration"] == 10136
    assert rec["signal"][:5].tolist() == [1129, 559, 551, 571, 576]
    assert rec["channel"]["sampling_rate"] == 3012.0 and rec["channel"]["digitisation"] == 8192.0


def test_extract_writes_signal_files(tmp_path):
    flags = types.SimpleNamespace(input_dir=os.path.join(GOLDEN, "fast5"), output_dir=str(tmp_path), mode="dna",
                                  unit=False, recursive=True, delimiter="\n", idname=False, threads=1, test_number=None)
    assert extract(flags) == 2
    out = chiron_input.read_signal(os.path.join(str(tmp_path), "raw", "read1.signal"))
    ref = chiron_input.read_signal(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    assert np.array_equal(out, ref)
    flags.mode, flags.output_dir = "rna", str(tmp_path / "rna")
    extract(flags)
    rna = chiron_input.read_signal(os.path.join(flags.output_dir, "raw", "rna_read_100_ch_328.signal"))
    assert rna[-5:].tolist() == [576, 571, 551, 559, 1129]          # rna mode reverses the signal (:165)
    assert os.path.isdir(os.path.join(flags.output_dir, "reference")) and os.path.isdir(flags.log_folder)


def test_normalisation_and_windows_match_oracle():
    sig = chiron_input.read_signal(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"))
    for mode in (NORM_UNIQUE_MAD, NORM_FULL_MAD):
        assert np.array_equal(chiron_input.normalize_signal(sig, mode), O.normalize_signal(sig, mode))
    norm = chiron_input.normalize_signal(sig, NORM_UNIQUE_MAD)
    for L, jump, start in ((400, 390, 0), (300, 100, 7), (512, 512, 0), (100000, 5, 62000)):
        ds = chiron_input.windows_from_signal(norm[start:], jump, L)
        x, lens = O.make_windows(norm, L, jump, start)
        assert np.array_equal(ds.event, x) and np.array_equal(ds.event_length, lens)
    ds = chiron_input.read_data_for_eval(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"), 0, 390, 400)
    assert ds.reads_n == 161
    seen = 0
    while ds.epochs_completed == 0:                                  # sequential, remainder at the end of the read
        xb, lb, _ = ds.next_batch(50)
        assert len(xb) == min(50, 161 - seen)
        seen += len(xb)
    assert seen == 161
    empty = chiron_input.windows_from_signal(norm[:0], 390, 400)
    assert empty.reads_n == 0


def test_presets_and_kernel_choice():
    a = types.SimpleNamespace(preset="dna-pre", mode="dna", start=None, batch_size=None, segment_len=None, jump=None,
                              threads=None, beam=None)
    a = apply_preset(a)
    assert (a.batch_size, a.segment_len, a.jump, a.beam, a.reverse_fast5) == (400, 400, 390, 30, False)
    b = apply_preset(types.SimpleNamespace(preset="rna-pre", mode="rna", start=None, batch_size=None, segment_len=None,
                                           jump=1000, threads=None, beam=0))
    assert (b.batch_size, b.segment_len, b.jump, b.beam, b.reverse_fast5) == (300, 2000, 1000, 0, True)
    with pytest.raises(ValueError):
        apply_preset(types.SimpleNamespace(preset="dna-pre", mode="rna", start=None, batch_size=None, segment_len=None,
                                           jump=None, threads=None, beam=None))
    assert [get_assembler_kernal(j, 300) for j in (270, 290, 300)] == ["simple", "glue", "stick"]
    files, d = list_input_files(types.SimpleNamespace(input=os.path.join(GOLDEN, "DNA", "raw"), recursive=True))
    assert files == ["read%d.signal" % i for i in range(1, 6)]


def test_recursive_listing_keeps_sub_folder_paths(tmp_path):
    """-r on a nested folder: the reference concatenates ``dirpath[dir_len:] + filename`` without a separator
    (chiron_eval.py:283) and then cannot open the file; here the input name keeps its relative path and the output
    prefix is flattened, so the read lands in the flat result/ segments/ meta/ folders."""
    from chiron_b200.chiron_eval import output_prefix
    (tmp_path / "sub" / "deep").mkdir(parents=True)
    for rel in ("a.signal", "sub/b.signal", "sub/deep/c.fast5", "sub/notes.txt"):
        (tmp_path / rel).write_text("1 2 3")
    files, d = list_input_files(types.SimpleNamespace(input=str(tmp_path), recursive=True))
    assert files == ["a.signal", os.path.join("sub", "b.signal"), os.path.join("sub", "deep", "c.fast5")]
    assert all(os.path.exists(os.path.join(d, f)) for f in files)
    assert [output_prefix(f) for f in files] == ["a", "sub__b", "sub__deep__c"]
    files, _ = list_input_files(types.SimpleNamespace(input=str(tmp_path), recursive=False))
    assert files == ["a.signal"]


def test_read_assignment_is_a_balanced_partition():
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 10 ** 6, size=57).tolist()
    for world in (1, 2, 4, 8):
        parts = shard.assign_reads(sizes, world)
        assert sorted(i for p in parts for i in p) == list(range(57))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)
    assert shard.assign_reads([], 4) == [[], [], [], []]


def _bcast_worker(rank, world, port, blob_path, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = open(blob_path, "rb").read() if rank == 0 else None
    got = shard.broadcast_blob(blob, src=0)
    files = ["r%d" % i for i in range(9)]
    mine = shard.assign_reads([9, 8, 7, 6, 5, 4, 3, 2, 1], world)[rank]
    import zlib
    q.put((rank, len(got), zlib.crc32(got), [files[i] for i in mine]))
    dist.destroy_process_group()


def test_two_rank_weight_broadcast_and_sharding_gloo(tmp_path):
    """The N>1 host logic on CPU: one broadcast of the weight blob, disjoint read shards, nothing else exchanged."""
    import torch.multiprocessing as mp
    from chiron_b200.model import bundled_blob_path
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, bundled_blob_path("DNA_default"), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res[0][1] == res[1][1] == os.path.getsize(bundled_blob_path("DNA_default"))
    assert res[0][2] == res[1][2]
    assert sorted(res[0][3] + res[1][3]) == ["r%d" % i for i in range(9)] and not set(res[0][3]) & set(res[1][3])


def test_host_pipeline_end_to_end_with_a_stub_gpu(tmp_path, monkeypatch):
    """chiron_eval.run() from files to files with the GPU replaced by a do-nothing Basecaller of the same surface
    (tools/call_bench.py's StubCaller): reader threads, cross-read batching, the two-slot submit/collect protocol, the
    finisher thread, the writer pool, the output tree and the JSON perf report -- everything but the kernels."""
    import json
    import shutil
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from call_bench import StubCaller
    from chiron_b200 import chiron_eval
    src = tmp_path / "in"
    src.mkdir()
    n_samples = {}
    for i in range(7):
        name = "read1.signal" if i % 2 == 0 else "read3.signal"
        shutil.copy(os.path.join(GOLDEN, "DNA", "raw", name), str(src / ("r%02d.signal" % i)))
        n_samples["r%02d" % i] = len(chiron_input.read_signal(os.path.join(GOLDEN, "DNA", "raw", name)))
    (src / "empty.signal").write_text("")                               # a read without samples still gets its files
    (src / "notes.txt").write_text("ignored")
    submitted = []

    reserved = []

    class Recorder(StubCaller):
        def basecall_submit(self, slot, x, seq_len, beam=0):
            submitted.append((slot, x.shape[0], int(seq_len.sum())))
            return super().basecall_submit(slot, x, seq_len, beam)

        def reserve_sms(self, n):              # cb_reserve_sms: asked for once, before the first batch, when a finisher runs
            reserved.append((n, len(submitted)))

    monkeypatch.setattr(chiron_eval, "Basecaller", lambda model, device=0, precision="auto": Recorder(model))
    monkeypatch.setenv("CHIRON_B200_GPU_BATCH", "300")
    out = str(tmp_path / "out")
    args = types.SimpleNamespace(input=str(src), output=out, model="DNA_default", start=None, batch_size=None, segment_len=None,
                                 jump=None, threads=3, beam=0, extension="fastq", concise=False, mode="dna", preset="dna-pre",
                                 precision="tc", recursive=False)
    chiron_eval.run(apply_preset(args))
    assert reserved == [(4, 0)]
    names = sorted(n_samples) + ["empty"]
    for sub in ("result", "segments"):
        assert sorted(os.listdir(os.path.join(out, sub))) == sorted(n + ".fastq" for n in names)
    assert {f for f in os.listdir(os.path.join(out, "meta"))} == {n + ".meta" for n in names} | {"all.meta", "all.perf.json"}
    with open(os.path.join(out, "meta", "all.perf.json")) as f:
        perf = json.load(f)
    assert perf["reads"] == 8 and perf["samples"] == sum(n_samples.values()) and perf["precision"] == "stub"
    assert perf["windows"] == sum(-(-n // 390) for n in n_samples.values())
    assert perf["Msamples_per_s"] > 0 and perf["world_size"] == 1 and perf["segment_len"] == 400 and perf["jump"] == 390
    # batching: windows are packed ACROSS reads into batches of >= 400 (the flag; the GPU batch was forced to 300), slots
    # alternate, every sample is submitted exactly once, only the last batch is partial
    assert [s[0] for s in submitted] == [i % 2 for i in range(len(submitted))]
    assert all(s[1] == 400 for s in submitted[:-1]) and 0 < submitted[-1][1] <= 400
    assert sum(s[1] for s in submitted) == perf["windows"]
    total_len = sum(min(400, n - st) for n in n_samples.values() for st in range(0, n, 390))
    assert sum(s[2] for s in submitted) == total_len
    # the stub calls 20 bases per window: every read's segment file has one record per window
    records = open(os.path.join(out, "segments", "r00.fastq")).read().split("\n")
    assert len([l for l in records if l.startswith(">r00")]) == -(-n_samples["r00"] // 390)
    assert open(os.path.join(out, "result", "empty.fastq")).read() == "@empty\n\n+\n\n"


def _sharded_call_worker(rank, world, port, src, out, q):
    """One rank of a read-sharded `chiron call` with the GPU stubbed (spawned: its own interpreter, like torchrun's)."""
    import sys
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from call_bench import StubCaller
    from chiron_b200 import chiron_eval
    chiron_eval.Basecaller = lambda model, device=0, precision="auto": StubCaller(model)
    args = types.SimpleNamespace(input=src, output=out, model="DNA_default", start=None, batch_size=None, segment_len=None,
                                 jump=None, threads=2, beam=0, extension="fasta", concise=True, mode="dna", preset="dna-pre",
                                 precision="auto", recursive=False)
    chiron_eval.run(apply_preset(args))
    q.put(rank)


def test_read_sharding_of_the_host_pipeline_with_a_stub_gpu(tmp_path):
    """Two ranks (separate processes with RANK / WORLD_SIZE / MASTER_* as torchrun sets them, gloo) basecall their own reads
    into ONE output folder: the result files partition the input, each rank leaves all.rank<r>.meta / .perf.json, and
    rank 0 merges them into meta/all.meta and meta/all.perf.json after one gather."""
    import json
    import multiprocessing as mp
    import shutil
    src = tmp_path / "in"
    src.mkdir()
    for i in range(9):
        shutil.copy(os.path.join(GOLDEN, "DNA", "raw", "read1.signal" if i % 3 else "read3.signal"), str(src / ("r%02d.signal" % i)))
    out = str(tmp_path / "out")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 777) % 2000
    procs = [ctx.Process(target=_sharded_call_worker, args=(r, 2, port, str(src), out, q)) for r in range(2)]
    for p in procs:
        p.start()
    assert sorted(q.get(timeout=180) for _ in range(2)) == [0, 1]
    for p in procs:
        p.join(60)
    assert {f[:-6] for f in os.listdir(os.path.join(out, "result"))} == {"r%02d" % i for i in range(9)}
    per_rank = []
    for rank in (0, 1):
        with open(os.path.join(out, "meta", "all.rank%d.perf.json" % rank)) as f:
            per_rank.append(json.load(f))
        assert per_rank[-1]["rank"] == rank and per_rank[-1]["world_size"] == 2
        assert os.path.exists(os.path.join(out, "meta", "all.rank%d.meta" % rank))
    assert 3 <= per_rank[0]["reads"] <= 6 and per_rank[0]["reads"] + per_rank[1]["reads"] == 9    # balanced by file size
    with open(os.path.join(out, "meta", "all.perf.json")) as f:
        merged = json.load(f)
    assert merged["reads"] == 9 and merged["samples"] == per_rank[0]["samples"] + per_rank[1]["samples"]
    assert merged["wall_s"] == max(r["wall_s"] for r in per_rank) and len(merged["per_rank"]) == 2
    lines = open(os.path.join(out, "meta", "all.meta")).read().split("\n")
    assert lines[0] == "# Wall_time Sys_time User_time Cpu_time" and len(lines[1].split()) == 4
    assert abs(float(lines[1].split()[0]) - merged["wall_s"]) < 2e-3


def test_batch_statistics_models_get_the_reference_batches(tmp_path, monkeypatch):
    """A model whose BatchNorm uses the moments of the current batch (HEAD's simple_global_bn) must see the reference's
    own batches: exactly -b windows each whatever CHIRON_B200_GPU_BATCH says, the last one wrap-padded to -b rows
    (chiron_eval.py:322-329,352-358); with population statistics the same run packs >= gpu_batch windows per batch."""
    import shutil
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from call_bench import StubCaller
    from chiron_b200 import _lib, chiron_eval
    src = tmp_path / "in"
    src.mkdir()
    for i in range(3):
        shutil.copy(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"), str(src / ("r%d.signal" % i)))
    n_win = 3 * 161
    for bn_mode, want_rows in ((_lib.BN_BATCH, [100] * 5), (_lib.BN_POPULATION, [300, n_win - 300])):
        seen = []

        class Recorder(StubCaller):
            def basecall_submit(self, slot, x, seq_len, beam=0):
                seen.append((x.copy(), seq_len.copy()))
                return super().basecall_submit(slot, x, seq_len, beam)

        monkeypatch.setattr(chiron_eval, "Basecaller", lambda model, device=0, precision="auto": Recorder(model, bn_mode))
        monkeypatch.setenv("CHIRON_B200_GPU_BATCH", "300")
        args = types.SimpleNamespace(input=str(src), output=str(tmp_path / ("out%d" % bn_mode)), model="DNA_default", start=0,
                                     batch_size=100, segment_len=400, jump=390, threads=2, beam=0, extension="fasta",
                                     concise=True, mode="dna", preset=None, precision="auto", recursive=False)
        summary = chiron_eval.evaluation(args)
        assert [len(x) for x, _ in seen] == want_rows
        assert sum(v["windows"] for v in summary.values()) == n_win
        if bn_mode == _lib.BN_BATCH:                      # 483 = 4 * 100 + 83: rows 83.. of the last batch repeat rows 0..
            x, ln = seen[-1]
            assert np.array_equal(x[83:], x[:17]) and np.array_equal(ln[83:], ln[:17])


def test_multi_read_and_single_read_fast5_layouts(tmp_path):
    """The multi-read layout ``read_*/Raw/Signal`` (extract_sig_ref.py:137-150,178-193) and the single-read layout with
    ``read_id`` / ``channel_id`` attributes, on files synthesised by tests/h5_writer.py (the bundled examples are all
    single-read without read_id): reader, extraction file names and separators, pA conversion, RNA reversal."""
    from h5_writer import write_multi_read_fast5, write_single_read_fast5
    from chiron_b200.utils.extract_sig_ref import extract, extract_file
    rng = np.random.default_rng(3)
    reads = {"read_%s" % k: (rng.integers(200, 900, size=n).astype(np.int16), rid)
             for k, n, rid in (("0a1b", 57, "id-a"), ("77ff", 64, None), ("c0de", 31, "id-c"))}
    src = tmp_path / "in"
    src.mkdir()
    write_multi_read_fast5(str(src / "batch0.fast5"), reads)
    chan = {"offset": np.float64(3.0), "range": np.float64(1400.5), "digitisation": np.float64(8192.0),
            "sampling_rate": np.float64(4000.0)}
    single = rng.integers(200, 900, size=45).astype(np.int16)
    write_single_read_fast5(str(src / "lone.fast5"), single, read_number=12, read_id="id-lone", channel=chan)

    got = fast5.read_fast5(str(src / "batch0.fast5"))
    assert [r["read_key"] for r in got] == sorted(reads)
    for r in got:
        assert np.array_equal(r["signal"], reads[r["read_key"]][0]) and r["read_id"] == reads[r["read_key"]][1]
    lone = fast5.read_fast5(str(src / "lone.fast5"))
    assert len(lone) == 1 and lone[0]["read_key"] == "Read_12" and lone[0]["read_id"] == "id-lone"
    assert lone[0]["channel"]["digitisation"] == 8192.0 and np.array_equal(fast5.read_raw_signal(str(src / "lone.fast5")), single)
    # unit=True: (raw + offset) * range / digitisation (extract_sig_ref.py:158-163); rna mode reverses
    (_, pa, _), = extract_file(str(src / "lone.fast5"), mode="dna", unit=True)
    np.testing.assert_allclose(pa, (single + 3.0) * 1400.5 / 8192.0)
    (_, rev, _), = extract_file(str(src / "lone.fast5"), mode="rna")
    assert np.array_equal(rev, single[::-1])

    flags = types.SimpleNamespace(input_dir=str(src), output_dir=str(tmp_path / "out"), unit=False, recursive=True, mode="dna",
                                  delimiter="\n", idname=False, threads=1, test_number=None)
    assert extract(flags) == 4
    raw = tmp_path / "out" / "raw"
    assert sorted(os.listdir(str(raw))) == sorted(["batch0%s.signal" % k for k in reads] + ["lone.signal"])
    for k, (sig, _) in reads.items():
        text = (raw / ("batch0%s.signal" % k)).read_text()
        assert text == " ".join(str(v) for v in sig.tolist())               # the multi-read branch joins with blanks
        assert np.array_equal(chiron_input.read_signal(str(raw / ("batch0%s.signal" % k))), sig.astype(np.float32))
    assert (raw / "lone.signal").read_text() == "\n".join(str(v) for v in single.tolist())
    flags.idname, flags.output_dir, flags.threads = True, str(tmp_path / "out2"), 2      # two spawned worker processes
    assert extract(flags) == 4                                              # read_id names where present, file names otherwise
    assert sorted(os.listdir(str(tmp_path / "out2" / "raw"))) == sorted(
        ["id-a.signal", "batch0read_77ff.signal", "id-c.signal", "id-lone.signal"])


def test_unreadable_input_fails_loudly_and_promptly(tmp_path, monkeypatch):
    """A signal file with a token that is not a number stops the run with the library's error (the reference's reader
    raises ValueError from float() at the same place) instead of hanging the reader / finisher / writer threads."""
    import shutil
    import sys
    import threading
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from call_bench import StubCaller
    from chiron_b200 import _lib, chiron_eval
    src = tmp_path / "in"
    src.mkdir()
    shutil.copy(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"), str(src / "a.signal"))
    (src / "b.signal").write_text("487 421 4x3 438\n")
    shutil.copy(os.path.join(GOLDEN, "DNA", "raw", "read1.signal"), str(src / "c.signal"))
    monkeypatch.setattr(chiron_eval, "Basecaller", lambda model, device=0, precision="auto": StubCaller(model))
    args = types.SimpleNamespace(input=str(src), output=str(tmp_path / "out"), model="DNA_default", start=None, batch_size=None,
                                 segment_len=None, jump=None, threads=2, beam=0, extension="fastq", concise=False, mode="dna",
                                 preset="dna-pre", precision="fp32", recursive=False)
    before = threading.active_count()
    with pytest.raises(_lib.ChironB200Error, match="4x3"):
        chiron_eval.run(apply_preset(args))
    deadline = time.time() + 10
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() <= before, "worker threads were left behind"


def test_chiron_call_cli_from_fast5_with_a_stub_gpu(tmp_path, monkeypatch):
    """`chiron call` (entry.main) end to end on the CPU: fast5 folder (a bundled single-read file and a synthesised multi-read
    file) -> raw/*.signal -> result / segments / meta, with the GPU stubbed out."""
    import shutil
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from call_bench import StubCaller
    from h5_writer import write_multi_read_fast5
    from chiron_b200 import chiron_eval, entry
    src = tmp_path / "in"
    src.mkdir()
    shutil.copy(os.path.join(GOLDEN, "fast5", "read1.fast5"), str(src / "read1.fast5"))
    rng = np.random.default_rng(4)
    write_multi_read_fast5(str(src / "multi.fast5"), {"read_%02d" % i: (rng.integers(300, 800, size=900 + 400 * i).astype(np.int16),
                                                                       "uuid-%d" % i) for i in range(3)})
    monkeypatch.setattr(chiron_eval, "Basecaller", lambda model, device=0, precision="auto": StubCaller(model))
    out = str(tmp_path / "out")
    entry.main(["call", "-i", str(src), "-o", out, "-p", "dna-pre", "--beam", "0", "-t", "2"])
    reads = ["multiread_00", "multiread_01", "multiread_02", "read1"]
    assert sorted(os.listdir(os.path.join(out, "raw"))) == [r + ".signal" for r in reads]
    assert sorted(os.listdir(os.path.join(out, "result"))) == [r + ".fastq" for r in reads]
    assert np.array_equal(chiron_input.read_signal(os.path.join(out, "raw", "read1.signal")),
                          chiron_input.read_signal(os.path.join(GOLDEN, "DNA", "raw", "read1.signal")))
    meta = open(os.path.join(out, "meta", "multiread_02.meta")).read().split("\n")
    assert meta[2] == "# read_len batch_size segment_len jump start_pos" and meta[3].split()[1:] == ["400", "400", "390", "0"]
    assert os.path.exists(os.path.join(out, "meta", "all.meta")) and os.path.exists(os.path.join(out, "log", "extract.log"))
